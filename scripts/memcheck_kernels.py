"""Selected kernel parity tests called directly (no pytest runner: it dies under the sanitizer before the first CUDA
call), for `compute-sanitizer --tool memcheck python scripts/memcheck_kernels.py`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
import test_gpu_kernels as T  # noqa: E402

T.test_conv3x3_non_power_of_two_planes(2, 96, 64, 64)
T.test_conv3x3_non_power_of_two_planes(2, 12, 128, 64)
T.test_conv3x3_non_power_of_two_planes(3, 8, 64, 64)
T.test_conv3x3(2, 8, 8, 128, 160)
T.test_gemm_split_k(77, 2560, 8192)
T.test_gemm(616, 768, 784)
T.test_gemm(1000, 136, 72)
T.test_gemm_epilogues_and_strides()
T.test_attention_fwd_bwd(2, 2, 256, 256, 40)
T.test_attention_fwd_bwd(1, 2, 300, 200, 80)
T.test_attention_fwd_bwd(2, 8, 1024, 77, 80)
T.test_attention_fwd_bwd(1, 2, 256, 77, 160)
T.test_attention_causal(2, 12, 77, 64)
T.test_groupnorm_fwd_bwd(2, 64, 64, True)
T.test_layernorm_fwd_bwd(130, 320, False)
T.test_layernorm_fwd_bwd(616, 768, True)
torch.cuda.synchronize()
print("memcheck kernels OK")
