"""fourierflow_b200 — B200-native (sm_100a CUDA) backend of fourierflow's Factorized-FNO forward.

Only the hot path lives here: csrc/ (CUDA kernels + C ABI, built into lib/libffno_b200.so) and the
host-side mirrors of the reference's operator modules / rollout routine.
"""
__version__ = "0.1.0"
