"""Error type of the public API (mirrors the reference's xmhw/exception.py:18)."""


class XmhwException(Exception):
    pass
