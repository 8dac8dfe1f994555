"""`DriverClientError` -- Python mirror of /root/reference/src/error.rs:6-32.

Every negative status of the C ABI (include/blaze_b200.h `bz_status`) maps onto one variant.
"""


class DriverClientError(Exception):
    variant = "Unknown"

    def __init__(self, message="", code=-8):
        super().__init__(message)
        self.code = code


class WriteError(DriverClientError):
    variant = "WriteError"


class ReadError(DriverClientError):
    variant = "ReadError"


class HBICAPNotReady(DriverClientError):
    variant = "HBICAPNotReady"


class InvalidPrimitiveParam(DriverClientError):
    variant = "InvalidPrimitiveParam"


class CsvError(DriverClientError):
    variant = "CsvError"


class LoadFailed(DriverClientError):
    variant = "LoadFailed"


class FileError(DriverClientError):
    variant = "FileError"


class Unknown(DriverClientError):
    variant = "Unknown"


class NoDevice(DriverClientError):
    """No usable CUDA device.  (The reference panics in `open_channel`, utils.rs:74.)"""
    variant = "NoDevice"


class NoResult(DriverClientError):
    """`wait_result`/`result` with an empty task queue (the reference would spin forever)."""
    variant = "NoResult"


_BY_CODE = {-1: WriteError, -2: ReadError, -3: HBICAPNotReady, -4: InvalidPrimitiveParam, -5: CsvError,
            -6: LoadFailed, -7: FileError, -8: Unknown, -9: NoDevice, -10: NoResult}


def check(rc):
    if rc == 0:
        return
    from ._lib import lib
    msg = lib().bz_last_error()
    raise _BY_CODE.get(rc, Unknown)((msg or b"").decode("utf-8", "replace"), rc)
