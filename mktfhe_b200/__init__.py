"""mktfhe_b200 -- B200-native (sm_100a) multi-key TFHE gate bootstrapping behind the SNUCP/MKTFHE API.

Host side (this package) mirrors /root/reference/src/MKTFHE.jl:21-35: parameter sets, `CRS` + `party_keygen`
(KeySet), `setup`, `lwe_encrypt` / `lwe_ith_encrypt` / `lwe_decrypt`, the six gates and `bootstrapping`.
The hot path runs in libmktfhe_b200.so (CUDA, include/mktfhe_b200.h); there is no CPU fallback.
"""
from . import params  # noqa: F401
from .keys import KeySet  # noqa: F401
from .params import ALL as PARAMS  # noqa: F401
from .scheme import MODE_FAST, MODE_STRICT, MktfheError, Scheme, setup  # noqa: F401
from .gate import AND, NAND, NOR, NOT, OR, XNOR, XOR, bootstrapping  # noqa: F401
