"""ctrlhair_b200 — B200-native (sm_100a) SEAN/SPADE generator hot path of XuyangGuo/CtrlHair.

Python here is plumbing (memory, streams, weight packing, the reference-facing call surface); the math runs in
the hand-written CUDA library ctrlhair_b200/lib/libctrlhair_b200.so behind the C ABI of include/ctrlhair_b200.h.
"""
__version__ = "0.1.0"
