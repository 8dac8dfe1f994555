"""`DriverClient`, `DriverConfig`, `CardType`, `DriverPrimitive` -- Python mirror of
/root/reference/src/driver_client/{dclient.rs,dclient_cfg.rs} over the C ABI.

Same names, argument meaning and error behaviour as the reference; the transport underneath is
CUDA (device arena + streams) instead of XDMA character devices.
"""
import ctypes
import enum

from ._lib import lib, buf_ptr
from .error import check


class CardType(enum.IntEnum):      # dclient_cfg.rs:1-3 (+ the card this build drives)
    C1100 = 0
    B200 = 1


class DriverConfig:
    """dclient_cfg.rs:9-47.  The FPGA address map has no meaning on a GPU; the object only
    carries the card type so that `DriverClient::new(id, DriverConfig::driver_client_cfg(..))`
    reads the same."""

    def __init__(self, card_type=CardType.B200):
        self.card_type = CardType(card_type)

    @staticmethod
    def driver_client_cfg(card_type=CardType.B200):
        return DriverConfig(card_type)


class DMA_RW(enum.IntEnum):        # dclient_code.rs:67-69
    OFFSET = 0


class DriverClient:
    """dclient.rs:50-86.  `id` is the slot string of the reference = CUDA device ordinal; a comma-separated list
    ("0,1,2,3") opens one client over several devices (MSMClient then shards over them)."""

    def __init__(self, id="0", cfg=None):
        self.cfg = cfg or DriverConfig()
        self.id = str(id)
        h = ctypes.c_void_p()
        check(lib().bz_dclient_new(self.id.encode(), int(self.cfg.card_type), ctypes.byref(h)))
        self._h = h

    @classmethod
    def new(cls, id, cfg):
        return cls(id, cfg)

    def close(self):
        if getattr(self, "_h", None):
            lib().bz_dclient_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- dclient.rs:88-93
    def reset(self):
        check(lib().bz_dclient_reset(self._h))

    # -- dclient.rs:500-517 / 456-471
    def dma_write(self, base_address, offset, data):
        p, n, keep = buf_ptr(data)
        check(lib().bz_dclient_dma_write(self._h, int(base_address), int(offset), p, n))

    def dma_read(self, base_address, offset, size):
        out = bytearray(size)
        p, n, keep = buf_ptr(out)
        check(lib().bz_dclient_dma_read(self._h, int(base_address), int(offset), p, n))
        return bytes(out)

    # -- FPGA-shell management surface (no-ops on a GPU; see include/blaze_b200.h)
    def firewalls_status(self):
        m = ctypes.c_uint32()
        check(lib().bz_dclient_firewalls_status(self._h, ctypes.byref(m)))
        return m.value

    def unblock_firewalls(self):
        check(lib().bz_dclient_unblock_firewalls(self._h))

    def initialize_cms(self):
        check(lib().bz_dclient_initialize_cms(self._h))

    def reset_sensor_data(self):
        check(lib().bz_dclient_reset_sensor_data(self._h))

    def setup_before_load_binary(self):
        check(lib().bz_dclient_setup_before_load_binary(self._h))

    def load_binary(self, binary):
        p, n, keep = buf_ptr(binary)
        check(lib().bz_dclient_load_binary(self._h, p, n))
        return 0

    # -- multi-GPU (B200 additions; see include/blaze_b200.h)
    def device_count(self):
        """Devices behind this client: 1, or the length of an id list such as "0,1,2,3"."""
        n = ctypes.c_uint32()
        check(lib().bz_dclient_device_count(self._h, ctypes.byref(n)))
        return n.value

    @staticmethod
    def comm_unique_id() -> bytes:
        """128-byte NCCL id: make it on one rank, hand it to every rank, pass it to comm_init."""
        out = ctypes.create_string_buffer(128)
        check(lib().bz_comm_unique_id(out))
        return out.raw

    def comm_init(self, rank: int, world: int, unique_id: bytes):
        """One process per GPU: this client becomes rank `rank` of `world`; MSM results are summed over the ranks
        device-side (NCCL all-gather + combine kernel on the client's stream)."""
        assert len(unique_id) == 128
        check(lib().bz_dclient_comm_init(self._h, int(rank), int(world), unique_id))

    def comm_info(self):
        r, w = ctypes.c_int32(), ctypes.c_int32()
        check(lib().bz_dclient_comm_info(self._h, ctypes.byref(r), ctypes.byref(w)))
        return r.value, w.value

    def device_info(self):
        name = ctypes.create_string_buffer(128)
        t, f = ctypes.c_uint64(), ctypes.c_uint64()
        check(lib().bz_dclient_device_info(self._h, name, 128, ctypes.byref(t), ctypes.byref(f)))
        return name.value.decode(), t.value, f.value


class DriverPrimitive:
    """dclient.rs:28-46: the 7-method trait every primitive client implements."""

    def loaded_binary_parameters(self):
        raise NotImplementedError

    def initialize(self, param):
        raise NotImplementedError

    def set_data(self, input):
        raise NotImplementedError

    def start_process(self, param=None):
        raise NotImplementedError

    def wait_result(self):
        raise NotImplementedError

    def result(self, param=None):
        raise NotImplementedError
