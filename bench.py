#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on B200: generated audio samples per second
(22.05 kHz) of batched MoL WaveRNN.generate(), BASELINE.json configs[1]:
a 10-second 80-bin mel (T=800 frames) folded with target=11000 / overlap=550 into 19 folds x
12,100 steps, rnn_dims=512, random-init weights, synthetic mel.

    python bench.py --gpus 1 --steps 5 --warmup 3            # this framework (CUDA engine)
    python bench.py --impl reference --steps 2 --warmup 1    # the reference algorithm on host cores
    torchrun ... bench.py --gpus N ...                       # N ranks, folds sharded, weak scaling

One "step" = one full pass of the hot path over the workload (all folds x all steps).
`value`  : fold-samples generated per second, inputs (upsampled conditioning, RNG draws)
           resident in HBM, timed with CUDA events on the launching stream, max over ranks.
`e2e`    : the same metric through the public call WaveRNN.generate(host mel, ...) including
           H2D of the mel and of the reference-compatible RNG draws, the conditioning network,
           the kernel, D2H of the samples and the host epilogue (xfade/unfold), wall clock.
N > 1    : weak scaling -- the mel is lengthened so every rank owns 19 folds; one NCCL
           all-gather of the sample blocks is inside the timed region.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FLOP_PER_SAMPLE_MOL = 7_650_304          # 3,825,152 MAC per fold-sample
BYTES_PER_SAMPLE = 836                   # 208 fp32 conditioning values in + 1 fp32 sample out
TARGET, OVERLAP, HOP = 11_000, 550, 275
FOLDS_PER_GPU = 19
CTOR = dict(rnn_dims=512, fc_dims=512, bits=9, pad=2, upsample_factors=(5, 5, 11), feat_dims=80,
            compute_dims=128, res_out_dims=128, res_blocks=10, hop_length=HOP, sample_rate=22050, mode='MOL')


def frames_for(n_gpus: int) -> int:
    """Mel frames giving exactly 19 folds per GPU with no padding (N=1 -> 800)."""
    L = FOLDS_PER_GPU * n_gpus * (TARGET + OVERLAP) + OVERLAP
    assert L % HOP == 0
    return L // HOP


def build_model(device, mode="MOL"):
    from wavernn_b200 import WaveRNN
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = WaveRNN(**dict(CTOR, mode=mode))
    m.gen_verbose = False
    return m.to(device)


def measured_traffic(args, model):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel on this workload, taken from
    the committed ncu --set full capture (profiles/r01_tc_traffic.json); None when no capture matches."""
    p = ROOT / "profiles" / "r01_tc_traffic.json"
    if not p.is_file() or args.workload != "cfg2" or args.seg_steps:
        return None
    d = json.loads(p.read_text())
    return d.get("dram_bytes_per_launch") if d.get("engine") == model.gen_stats.get("engine") else None


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.is_file():
        d = json.loads(p.read_text())
        return dict(tflops=float(d["bf16_tflops"]), tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    hbm_gbs=float(d["hbm_gbs"]), src="measured")
    return dict(tflops=1590.0, tflops_sustained=1400.0, hbm_gbs=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port on torch CPU operators) on the host cores
# ------------------------------------------------------------------------------------------
def cpu_conditioning(model_cpu, mel):
    with torch.no_grad():
        model_cpu.eval()
        mp = torch.nn.functional.pad(mel, (2, 2))
        m_up, aux = model_cpu.upsample(mp)
        mels_f = model_cpu.fold_with_overlap(m_up, TARGET, OVERLAP)
        aux_f = model_cpu.fold_with_overlap(aux, TARGET, OVERLAP)
    return mels_f, aux_f


def time_cpu_port(sample_steps: int, repeats: int = 1, warmup: int = 0):
    from oracle import torch_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = build_model("cpu")
    torch.manual_seed(0)
    mel = torch.rand(1, 80, frames_for(1))
    mels_f, aux_f = cpu_conditioning(model, mel)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    B = mels_f.shape[0]
    # The loop is 19-row GEMVs: on a many-core host torch's default (all cores) oversubscribes badly
    # (128 threads: ~40x slower than 8).  Give the CPU arm its best thread count from a short sweep.
    trials = {}
    for n in sorted({min(cores, c) for c in (4, 8, 16, 32, cores)}):
        torch.set_num_threads(n)
        torch.manual_seed(1234)
        _, dt = torch_port.generate_segments_torch(sd, mels_f, aux_f, steps=40)
        trials[n] = dt
    # the 40-step sweep is noisy on a shared host: time the full sample on the two best counts, keep the faster
    finals = {}
    for n in sorted(trials, key=trials.get)[:2]:
        torch.set_num_threads(n)
        ts = []
        for i in range(warmup + repeats):
            torch.manual_seed(1234)
            _, dt = torch_port.generate_segments_torch(sd, mels_f, aux_f, steps=sample_steps)
            if i >= warmup:
                ts.append(dt)
        finals[n] = ts
    best = min(finals, key=lambda n: float(np.mean(finals[n])))
    torch.set_num_threads(best)
    cores, times = best, finals[best]
    return dict(B=B, steps=sample_steps, seconds=times, cores=cores, threads=torch.get_num_threads(),
                host_cores=os.cpu_count() or 1, sweep={str(k): round(40 * B / v, 1) for k, v in trials.items()})


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_steps = 1210                          # 10% of the 12,100 steps of every fold, all 19 folds
    r = time_cpu_port(sample_steps, repeats=args.steps, warmup=args.warmup)
    per_step = float(np.mean(r["seconds"]))
    value = r["B"] * sample_steps / per_step
    sample = f"{r['B']} folds x {sample_steps} of 12100 steps per bench step (torch CPU operators, best of thread sweep = {r['threads']} threads on {r['host_cores']} host cores; samples/s by threads: {r['sweep']})"
    line = {"impl": "reference", "metric": "audio samples/sec (22.05 kHz) batched MoL generate", "value": value,
            "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2: 10 s mel (T=800) -> 19 folds x 12100 steps, target=11000 overlap=550, "
                                   "MoL head, rnn_dims=512, random-init", "sample": sample},
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": r["cores"], "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "x_realtime": value / 22050.0}
    emit(line)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from wavernn_b200.sharding import fold_geometry, gather_segments, shard_folds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    if args.workload == "cfg4":
        return run_cfg4(args, device, rank, world, local)
    cfg3 = args.workload == "cfg3"                     # BASELINE configs[2]: cfg2's mel with the 9-bit RAW head (mu-law)
    model = build_model(device, "RAW" if cfg3 else "MOL")
    model.gen_precision, model.gen_engine = args.precision, args.engine
    cfg5 = args.workload == "cfg5"
    if cfg3:
        model.gen_rng = "philox"                       # (parity mode would stream 471 MB of Exp(1) draws per call)
    if cfg5:
        # BASELINE configs[4] (SURVEY 8d choice A): T=172,034 frames = 35.8 min -> exactly 4096 folds with the
        # reference's own target/overlap; the 4096 folds are sharded over the ranks (strong scaling); in-kernel RNG
        model.gen_rng = "philox"
    T = 172_034 if cfg5 else frames_for(world)
    torch.manual_seed(0)
    mel_host = torch.rand(1, 80, T).pin_memory()
    geo = fold_geometry(T * HOP, TARGET, OVERLAP)
    shard = shard_folds(geo, rank, world, HOP)
    B_total, S = geo.n_seg, geo.seg_len

    # ---- resident inputs for the device-timed region ------------------------------------
    model.eval()
    with torch.no_grad():
        mp = torch.nn.functional.pad(mel_host.to(device), (2, 2))
        m_up, aux = model.conditioning(mp, shard.frame_lo, shard.frame_hi)
        off = shard.row_lo - shard.frame_lo * HOP
        m_up = m_up[off:off + shard.row_hi - shard.row_lo].contiguous()
        aux = aux[off:off + shard.row_hi - shard.row_lo].contiguous()
    f0, n = shard.seg_first, shard.n_seg
    if cfg5 or cfg3:
        u_all, uni_ptr, h2d_rng = None, 0, 0
    else:
        torch.manual_seed(1234)
        u_all, _ = model._reference_draws(geo, S)
        uni = torch.cat([u_all[:, 10 * f0:10 * (f0 + n)], u_all[:, 10 * B_total + f0:10 * B_total + f0 + n]], 1).contiguous().to(device)
        uni_ptr, h2d_rng = uni.data_ptr(), u_all.numel() * 4
    out = torch.empty((n, S), dtype=torch.float32, device=device)
    engine = model._get_engine(device)
    stream = torch.cuda.current_stream(device)

    def device_step():
        engine.generate(mels_up=m_up.data_ptr(), aux=aux.data_ptr(), L=m_up.shape[0], n_seg=n, seg_len=S,
                        seg_stride=geo.seg_stride, out=out.data_ptr(), seg_first=f0, uniforms=uni_ptr,
                        steps=args.seg_steps, stream=stream.cuda_stream)
        if world > 1:
            return gather_segments(out, shard, geo)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(args.warmup):
        device_step()
    barrier(); engine.check()
    launches0 = engine.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with ClockSampler(local) as clk:
        barrier()
        ev[0].record(stream)
        for _ in range(args.steps):
            device_step()
        ev[1].record(stream)
        barrier()
    engine.check()
    t_dev = ev[0].elapsed_time(ev[1]) / 1e3
    gpu_launches = engine.launch_count - launches0
    clocks = clk.summary()

    # ---- end to end through the public API, host buffers --------------------------------
    def e2e_step():
        torch.manual_seed(1234)
        return model.generate(mel_host, None, True, TARGET, OVERLAP, cfg3)

    if args.skip_e2e:
        t_e2e = float("nan")
    else:
        for _ in range(min(args.warmup, 2)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            wav = e2e_step()
        barrier()
        t_e2e = time.perf_counter() - t0
        assert np.isfinite(wav).all()

    times = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = times.tolist()

    if rank == 0:
        S_eff = args.seg_steps or S
        units = B_total * S_eff * args.steps
        value, e2e_value = units / t_dev, units / t_e2e
        peaks = measured_peaks()
        long_run = (t_dev / args.steps) > 1.0
        peak_tf = peaks["tflops_sustained"] if long_run else peaks["tflops"]
        ach_tf = value / world * (8_143_872 if cfg3 else FLOP_PER_SAMPLE_MOL) / 1e12          # per GPU
        step_us = t_dev / args.steps / S_eff * 1e6
        cpu = None
        if world == 1 and not args.no_cpu_baseline and not cfg3:
            r = time_cpu_port(args.cpu_sample_steps)
            v = r["B"] * r["steps"] / r["seconds"][0]
            cpu = {"value": v, "unit": "samples/s", "cores": r["cores"], "kind": "port",
                   "sample": f"{r['B']} folds x first {r['steps']} of 12100 steps, torch CPU operators, best of thread sweep = {r['threads']} threads on {r['host_cores']} host cores; samples/s by threads: {r['sweep']}"}
        line = {
            "metric": "audio samples/sec (22.05 kHz) batched " + ("RAW 9-bit" if cfg3 else "MoL") + " generate", "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong" if cfg5 else "weak", "vs_baseline": None,
            "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": (f"cfg5: mel T={T} frames (35.8 min) -> {B_total} folds x {S} steps sharded over {world} GPU(s), "
                                    if cfg5 else
                                    f"cfg2 x{world}: mel T={T} frames -> {B_total} folds x {S} steps ({FOLDS_PER_GPU} folds per GPU), ")
                                   + f"target={TARGET} overlap={OVERLAP}, " + ("RAW head (bits=9, mu-law)" if cfg3 else "MoL head") + ", rnn_dims=512, random-init weights, torch.rand mel",
                       "engine": model.gen_stats.get("engine", engine.name), "grid_ctas": engine.grid_ctas,
                       "parallelism": f"folds sharded x{world}" if world > 1 else "single GPU",
                       "l2_policy": "no flush: the per-step conditioning stream (183 MB per 19 folds) exceeds the 126 MB L2",
                       "rng": "in-kernel Philox4x32-10" if (cfg5 or cfg3) else "reference-compatible torch CPU draws, resident in HBM for `value`"},
            "clocks": clocks, "gpu_launches": int(gpu_launches),
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(mel_host.numel() * 4 + h2d_rng),
                    "d2h_bytes_per_step": int(B_total * S * 4), "ms_per_step": t_e2e / args.steps * 1e3},
            "roofline": {"bound": "tensor", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": ach_tf / peak_tf, "traffic": measured_traffic(args, model), "peak_source": peaks["src"],
                         "note": "latency/sync-bound at 19 folds per GPU (SURVEY 8d): fraction of the dense-fp16 "
                                 "tensor roofline is reported for completeness",
                         "hbm_achieved_gbs": value / world * BYTES_PER_SAMPLE / 1e9, "hbm_peak_gbs": peaks["hbm_gbs"]},
            "x_realtime": value / 22050.0, "x_realtime_per_fold": 1.0 / (step_us * 1e-6) / 22050.0,
            "us_per_sequential_step": step_us,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if args.seg_steps or args.skip_e2e:
            line["partial"] = "profiling run (--seg-steps/--skip-e2e): NOT a bench value"
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_cfg4(args, device, rank, world, local):
    """BASELINE configs[3] without the text front-end ('16 synthetic mels of 150-800 frames'): the folds of
    16 utterances vocoded as ONE job through WaveRNN.generate_many (sharded over the ranks, one all-gather).  Tacotron
    itself is out of scope and stays torch.  The whole call is host-facing, so `value` and `e2e` are the same
    end-to-end measurement (host mels in, float64 waveforms out), stated in `config`."""
    import torch.distributed as dist
    model = build_model(device)
    model.gen_precision, model.gen_engine, model.gen_rng = args.precision, args.engine, "philox"
    rs = np.random.RandomState(0)
    frames = [int(t) for t in rs.randint(150, 801, size=16)]
    torch.manual_seed(0)
    mels = [torch.rand(1, 80, T).pin_memory() for T in frames]
    from wavernn_b200.sharding import fold_geometry
    S = TARGET + 2 * OVERLAP
    folds = [fold_geometry(T * HOP, TARGET, OVERLAP).n_seg for T in frames]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def step():
        return model.generate_many(mels, [None] * len(mels), TARGET, OVERLAP, False)

    for _ in range(args.warmup):
        step()
    engine = model._get_engine(device)
    launches0 = engine.launch_count
    with ClockSampler(local) as clk:
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            wavs = step()
        barrier()
        t = time.perf_counter() - t0
    assert all(np.isfinite(w).all() for w in wavs)
    tt = torch.tensor([t], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t = float(tt.item())
    if rank == 0:
        units = sum(folds) * S * args.steps
        value = units / t
        peaks = measured_peaks()
        ach_tf = value / world * FLOP_PER_SAMPLE_MOL / 1e12
        emit({"metric": "audio samples/sec (22.05 kHz) batched MoL generate", "value": value, "unit": "samples/s",
              "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3,
              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
              "config": {"workload": f"cfg4: 16 synthetic utterances of {min(frames)}-{max(frames)} mel frames ({sum(frames)} frames, "
                                     f"{sum(frames) * HOP / 22050:.1f} s of audio) -> {sum(folds)} folds x {S} steps in ONE generate_many job, "
                                     f"sharded over {world} GPU(s); Tacotron not included (torch, out of scope)",
                         "engine": model.gen_stats.get("engine"), "rng": "in-kernel Philox4x32-10",
                         "note": "value == e2e: the job is timed end to end through the public call (host mels in, float64 waveforms out)"},
              "clocks": clk.summary(), "gpu_launches": int(engine.launch_count - launches0),
              "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": int(sum(frames) * 80 * 4),
                      "d2h_bytes_per_step": int(sum((T - 1) * HOP for T in frames) * 8), "ms_per_step": t / args.steps * 1e3},
              "roofline": {"bound": "tensor", "achieved": ach_tf, "peak": peaks["tflops"], "unit": "TFLOP/s",
                           "frac": ach_tf / peaks["tflops"], "traffic": None, "peak_source": peaks["src"]},
              "x_realtime": value / 22050.0, "delivered_audio_x_realtime": sum((T - 1) * HOP for T in frames) * args.steps / t / 22050.0})
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def guard_stdout():
    """The contract is ONE JSON line on stdout.  Libraries below us write there too (NCCL prints its version banner
    to fd 1 whenever NCCL_DEBUG is VERSION/WARN/INFO in the environment): keep a private handle on the real stdout
    for the JSON line and point fd 1 at stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tcgen05", "stream"])
    ap.add_argument("--cpu-sample-steps", type=int, default=3000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"],
                    help="cfg2 (default, the headline config: 19 folds per GPU) or cfg5 (4096 folds of a 35.8-min mel, "
                         "sharded over the ranks, in-kernel Philox draws)")
    ap.add_argument("--seg-steps", type=int, default=0,
                    help="PROFILING ONLY: generate just the first N steps of every fold in the device-timed region "
                         "(keeps ncu captures short); the printed line is then marked partial and is not a bench value")
    ap.add_argument("--skip-e2e", action="store_true", help="PROFILING ONLY: skip the end-to-end leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                             "(use --impl reference for the CPU arm)")
        run_ours(args)


if __name__ == "__main__":
    main()
