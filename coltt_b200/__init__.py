"""coltt_b200 — B200-native ANN search path behind coltt's edge Vectorstore/Quantization and
core/vectorindex HNSW interfaces.

The product is the C-ABI shared library (include/coltt_b200.h, coltt_b200/csrc -> coltt_b200/lib);
this package is the Python host-side mirror of the reference's Go interfaces used by tests and
bench.py (the Go toolchain is absent from the build image; INTEGRATION.md holds the cgo shim).
There is no CPU fallback: importing works anywhere, every compute call needs an sm_100 GPU.
"""
from ._lib import ColttError, lib, build_library, LIB_PATH, pinned_empty  # noqa: F401
from .edge import (  # noqa: F401
    Vectorstore, VectorSpace, SearchResultItem, Metadata,
    Distance_Cosine, Distance_Euclidean,
    Quantization_None, Quantization_F16, Quantization_F8, Quantization_BF16, Quantization_F8_E4M3,
    SELECT_COMPAT, SELECT_NEAREST, MATH_EXACT, MATH_FAST, score_helper,
)
from .vectorindex import Hnsw  # noqa: F401,E402
from .experimental import MultiVectorVertex, MultiVectorIndex, NearestNeighbor  # noqa: F401,E402
from .batcher import MicroBatcher  # noqa: F401,E402
