"""dformer_b200 -- B200-native (sm_100a) implementation of the DFormer RGB-D forward/backward hot path.

Drop-in mirror of the reference model API (`models/builder.py: EncoderDecoder`, the `DFormer_*`
constructors, `LightHamHead`; identical state_dict layout) on top of hand-written CUDA kernels
reached through the C ABI in include/dfb200.h.  Importing the package loads the shared library and
raises if it is missing -- there is no CPU or eager-PyTorch fallback."""
from ._lib import lib as _load

_load()          # fail loudly, at import, when libdformer_b200.so has not been built

from .models.builder import EncoderDecoder  # noqa: E402,F401
from .models.encoders.DFormer import DFormer, DFormer_Base, DFormer_Large, DFormer_Small, DFormer_Tiny  # noqa: E402,F401
from .models.decoders.ham_head import LightHamHead  # noqa: E402,F401
