"""protocols.secure_comparison_b200 -- B200-native batched engine for the DGK/Veugen secure comparison.

Drop-in surface: the reference's ``Initiator`` / ``KeyHolder`` ``step_1 ... step_7`` API
(``tno.mpc.protocols.secure_comparison`` 4.4.0) over batches of comparisons; the arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI of ``include/scb200.h``."""
from .engine import Engine, EngineError, NotInvertibleError  # noqa: F401
from .ciphertexts import BatchedCiphertext, BatchedDGKCiphertext, BatchedPaillierCiphertext  # noqa: F401
from .communicator import Communicator, InMemoryCommunicator  # noqa: F401
from .initiator import Initiator  # noqa: F401
from .keyholder import KeyHolder  # noqa: F401
from .schemes import SchemeEngine, as_limbs, get_engine, reset_engines, set_default_device  # noqa: F401
from .utils import from_bits, to_bits  # noqa: F401
from .limbs import int_to_limbs, ints_to_limbs, limbs_for_bits, limbs_to_int, limbs_to_ints  # noqa: F401

__version__ = "0.2.0"
