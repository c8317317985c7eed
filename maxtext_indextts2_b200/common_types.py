"""Constants shared by the decode path.

Mirrors the numerics-relevant constants of the reference
(MaxText/common_types.py:62-74): the model-mode strings, the active-sequence
indicator used in KV-cache segment ids, and the mask value applied to
attention scores.
"""

import numpy as np

MODEL_MODE_AUTOREGRESSIVE = "autoregressive"
MODEL_MODE_PREFILL = "prefill"
MODEL_MODE_TRAIN = "train"

# MaxText/common_types.py:70
DECODING_ACTIVE_SEQUENCE_INDICATOR = 1

# MaxText/common_types.py:74
DEFAULT_MASK_VALUE = -0.7 * float(np.finfo(np.dtype("float32")).max)

# MaxText/inference_utils.py:20 -- masking value used by nucleus sampling.
NEG_INF = -1.0e7
