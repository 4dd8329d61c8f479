"""gill_b200 -- B200-native (sm_100a) implementation of GILL's image-emission hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); every heavy operation is a hand-written
CUDA kernel behind the C ABI of libgillb200.so (include/gillb200.h).
"""
__version__ = "0.1.0"
