"""forgex_b200 -- B200-native implementation of Forgex's DFA matching hot path.

Only what the path needs lives here: csrc/ (host-side pattern compiler, sm_100a kernels, C ABI),
the ctypes binding and the host-side mirror of the reference's interface.
"""
from .api import (INVALID_CHAR_INDEX, ForgexError, Pattern, RegexResult, is_valid_regex, is_valid_regex_batch, launch_count, op_in,  # noqa
                  op_match, regex, regex_f, status_message)
