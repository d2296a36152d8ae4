"""Stand-in for the `torch_scatter` package, which the reference imports at module top
(lib/dvgo.py:10 `from torch_scatter import segment_coo`, lib/dmpigo.py:11 also `scatter_add`) but
does not pin or vendor.  `directvoxgo_b200.dropin.install()` registers this module under the name
`torch_scatter` when the real package is absent.  Only what DirectVoxGO calls is provided."""
from ..ops import scatter_add, segment_coo

__all__ = ["segment_coo", "scatter_add"]
