"""B200-native (sm_100a) cross-modal (query, product) match scorer.

Drop-in for the scoring hot path of zuokai/KDDCUP_2020_MultimodalitiesRecall_2nd_Place: the ImageBert (zk, lds)
and LXMERT encoders + match heads + the ensemble of code/main.py.  Host code is Python/PyTorch (device memory,
streams, torch.distributed); all arithmetic runs in hand-written CUDA kernels behind the C ABI of
include/mmrecall.h (libmmrecall.so, built in-tree by csrc/build.py).  There is no CPU fallback.
"""
__version__ = "0.1.0"
