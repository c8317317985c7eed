"""B200-native decode step for the IndexTTS2-on-MaxText text+audio-token transformer.

Host side mirrors the reference's engine API (``MaxEngine.prefill / insert /
generate``, MaxText/maxengine.py) and config keys (MaxText/configs/base.yml); the
arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI declared in
``include/mtx_b200.h``.  There is no CPU path: importing the compute modules without
the built library raises.
"""

__version__ = "0.1.0"
