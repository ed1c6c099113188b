# Default hyper-parameters for the vocoder path -- same names and values as the
# reference's hparams.py:20-60 (the names are the interface `hp.configure()` reads).
# Only the entries the WaveRNN.generate() path and its callers consume are listed;
# point `hp.configure()` at the reference's own hparams.py for the full set.

# DSP
sample_rate = 22050
n_fft = 2048
fft_bins = n_fft // 2 + 1
num_mels = 80
hop_length = 275
win_length = 1100
fmin = 40
min_level_db = -100
ref_level_db = 20
bits = 9
mu_law = True
peak_norm = False

# vocoder model
voc_model_id = 'ljspeech_mol'
voc_mode = 'MOL'
voc_upsample_factors = (5, 5, 11)
voc_rnn_dims = 512
voc_fc_dims = 512
voc_compute_dims = 128
voc_res_out_dims = 128
voc_res_blocks = 10
voc_pad = 2

# generation
voc_gen_batched = True
voc_target = 11_000
voc_overlap = 550
voc_gen_at_checkpoint = 5
