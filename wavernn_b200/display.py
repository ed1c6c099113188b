"""Progress output used by generate() (reference utils/display.py:9-18,21-81)."""
import sys
import time  # noqa: F401  (the reference re-exports `time` through its star import)


def progbar(i, n, size=16):
    done = (i * size) // n
    return ''.join('█' if j <= done else '░' for j in range(size))


def stream(message):
    sys.stdout.write(f"\r{message}")


def simple_table(item_tuples):
    cols = []
    for head, cell in item_tuples:
        head, cell = str(head), str(cell)
        width = max(len(head), len(cell))
        cols.append((head.center(width), cell.center(width)))
    border = '+' + '+'.join('-' * (len(h) + 2) for h, _ in cols) + '+'
    print(border)
    print('|' + '|'.join(f' {h} ' for h, _ in cols) + '|')
    print(border)
    print('|' + '|'.join(f' {c} ' for _, c in cols) + '|')
    print(border)
    print(' ')
