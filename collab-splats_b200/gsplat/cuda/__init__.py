"""CUDA operator layer (ctypes -> librade_b200.so); mirrors the module path ``gsplat.cuda._wrapper``."""
