"""vsim_b200 -- B200-native back-end for vsim's Q4_0/Q4_1 mul_mat hot path.

The product is ``libb200q4.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/b200q4.h`` plus the ``imax_*`` hook of ``csrc/imax_b200.c``).  This
package only holds the build recipe and a thin ctypes binding used by the
tests, ``bench.py`` and the multi-GPU harness.  There is no CPU fallback: every
compute entry point raises :class:`B200Q4Error` when no sm_100 device is usable.
"""
from .lib import (B200Q4Error, Chain, Q4_0, Q4_1, Q4Weight, device_info, init, is_initialised, kernel_for,
                  launch_count, launch_count_reset, load_library, mul_mat, mul_mat_transposed,
                  mul_mat_wdata, quantize_rows, quantize_weights, row_bytes, shutdown)

__all__ = ["B200Q4Error", "Chain", "Q4_0", "Q4_1", "Q4Weight", "device_info", "init", "is_initialised",
           "kernel_for", "launch_count", "launch_count_reset", "load_library", "mul_mat",
           "mul_mat_transposed", "mul_mat_wdata", "quantize_rows", "quantize_weights", "row_bytes", "shutdown"]
