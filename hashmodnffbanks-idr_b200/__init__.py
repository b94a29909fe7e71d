"""idrk - B200-native (sm_100a) implementation of the IDR hash-grid rendering hot path.

Host side: Python modules with the reference's names, constructor signatures and state-dict
keys (model/..., utils/rend_util.py).  Device side: hand-written CUDA in csrc/ behind the C ABI
declared in include/idrk.h, loaded with ctypes by `_lib`.  There is no CPU or PyTorch fallback:
calling a kernel without libidrk.so or without a CUDA device raises.
"""
__version__ = "0.1.0"
