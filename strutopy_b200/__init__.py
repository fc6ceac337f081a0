"""strutopy_b200 — B200-native variational-EM core for the Structural Topic Model, behind the
`STM(...)` / `expectation_maximization()` / `E_step()` / `M_step()` surface of mkrcke/strutopy
(/root/reference/src/modules/stm.py:310).  Hand-written CUDA for sm_100a through a C ABI
(include/stm_b200.h); no CPU fallback."""
from . import _lib  # noqa: F401
from .corpus import pack_corpus, read_mm, write_mm  # noqa: F401
from .generate_docs import CorpusCreation, sample_corpus  # noqa: F401
from .heldout import cut_in_half, eval_heldout, split_corpus  # noqa: F401
from .spectral import spectral_init  # noqa: F401
from .stm import STM  # noqa: F401

__all__ = ["STM", "pack_corpus", "read_mm", "write_mm", "eval_heldout", "cut_in_half", "split_corpus",
           "spectral_init", "CorpusCreation", "sample_corpus"]
