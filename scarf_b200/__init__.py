"""scarf_b200 -- B200-native implementation of Scarf's cell-graph construction path
(``DataStore.make_graph`` / ``mark_hvgs`` / ``run_mapping``).

Host code is Python + PyTorch (allocation, streams, torch.distributed); all arithmetic on the
path runs in hand-written sm_100a CUDA behind the C-ABI of ``include/scarf_b200.h``.  There is
no CPU fallback: importing :mod:`scarf_b200.lib` raises if the CUDA library has not been built.
"""
__version__ = "0.1.0"
