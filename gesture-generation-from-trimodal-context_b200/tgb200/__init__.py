"""tgb200 - B200-native kernels and launch plans behind the trimodal-gesture nn.Module API."""
from . import _lib  # noqa: F401
