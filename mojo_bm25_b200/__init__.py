"""mojo_bm25_b200 -- B200-native (sm_100a) BM25 query hot path behind the reference's Python API.

Only the hot path lives here (CSC posting gather -> score accumulation -> top-k) plus the thin
host-side mirrors of the reference's retrieval interfaces.  All numerics run in
libbm25_b200.so (hand-written CUDA); there is no CPU fallback.
"""
__version__ = "0.1.0"
