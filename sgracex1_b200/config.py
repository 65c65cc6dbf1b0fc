"""The reference's flag module (demo/*/config.py:3-35): same names, same defaults, except
`acc = 1` (the accelerator is what this package is).  The driver mirror `sgracex1_b200.sgrace`
reads these globals at every forward, as the reference does."""
import numpy as np

device = "cpu"
hidden_channels = 16
layer_count = 1          # layers processed per hardware call
load_weights = 1

accb = 0                 # accelerator in the backward path (the mirror keeps the saved-tensor backward; gemm_mode 2 runs in float32 mode)
acc = 1                  # accelerator in the forward path
show_max_min = 0
min_output = 1
profiling = 0
fake_quantization = 1
hardware_quantize = 1
compute_attention = 0    # GAT at 1, GCN at 0
stream_mode = 0
head_count = 1

N_adj = 20480            # buffer maxima used by init_SGRACE
M_adj = 20480
M_fea = 2048
P_w = hidden_channels
NNZ_adj = 1000000
NNZ_fea = 4000000
w_qbits = 8

float_type = np.float32
