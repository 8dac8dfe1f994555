"""`NTTClient` -- Python mirror of /root/reference/src/ingo_ntt/ntt_api.rs over the C ABI.

    NTTClient.new(NTT.Ntt, dclient)                              ntt_api.rs:26-31 (2^27, BLS12-381 Fr)
    set_data(NTTInput{buf_host, data}) / initialize(NttInit{}) / start_process(Some(buf_kernel))
    / wait_result() / result(Some(buf))                          ntt_api.rs:37-124
`NTTClient.new_ex(...)` is the B200 addition for other sizes / fields / the inverse transform.
"""
import ctypes
import enum
from dataclasses import dataclass

from ._lib import lib, buf_ptr
from .driver_client import DriverClient, DriverPrimitive
from .error import check


class NTT(enum.IntEnum):       # ntt_api.rs:8-10
    Ntt = 0


@dataclass
class NttInit:                 # ntt_api.rs:17
    pass


@dataclass
class NTTInput:                # ntt_api.rs:19-23
    buf_host: int
    data: object


class NTTClient(DriverPrimitive):
    NTT_LOG_SIZE = 27          # ntt_data.rs:65-66

    def __init__(self, ptype=NTT.Ntt, dclient: DriverClient = None, *, field=2, log_size=None, inverse=False):
        self.driver_client = dclient
        h = ctypes.c_void_p()
        if log_size is None:
            check(lib().bz_ntt_new(dclient._h, int(ptype), ctypes.byref(h)))
            log_size = self.NTT_LOG_SIZE
        else:
            check(lib().bz_ntt_new_ex(dclient._h, int(field), int(log_size), 1 if inverse else 0, ctypes.byref(h)))
        self._h = h
        self.log_size = log_size
        self.nbytes = (1 << log_size) * 32

    @classmethod
    def new(cls, ptype, dclient):
        return cls(ptype, dclient)

    @classmethod
    def new_ex(cls, dclient, field=2, log_size=27, inverse=False):
        return cls(NTT.Ntt, dclient, field=field, log_size=log_size, inverse=inverse)

    def close(self):
        if getattr(self, "_h", None):
            lib().bz_ntt_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def loaded_binary_parameters(self):                       # todo!() in the reference (ntt_api.rs:33-35)
        out = (ctypes.c_uint32 * 2)()
        check(lib().bz_ntt_loaded_binary_parameters(self._h, out))
        return [out[0], out[1]]

    def initialize(self, param=None):                         # ntt_api.rs:37-56
        check(lib().bz_ntt_initialize(self._h))

    def set_data(self, input: NTTInput):                      # ntt_api.rs:72-87
        p, n, keep = buf_ptr(input.data)
        check(lib().bz_ntt_set_data(self._h, int(input.buf_host), p, n))

    def start_process(self, buf_kernel=None):                 # ntt_api.rs:58-70 (unwrap()s None)
        if buf_kernel is None:
            from .error import InvalidPrimitiveParam
            raise InvalidPrimitiveParam("start_process needs Some(buf_kernel)", -4)
        check(lib().bz_ntt_start_process(self._h, int(buf_kernel)))

    def wait_result(self):                                    # ntt_api.rs:89-108
        check(lib().bz_ntt_wait_result(self._h))

    def result(self, buf_num=None, out=None):                 # ntt_api.rs:110-124
        if buf_num is None:
            from .error import InvalidPrimitiveParam
            raise InvalidPrimitiveParam("result needs Some(buf_num)", -4)
        if out is None:
            out = bytearray(self.nbytes)
        p, n, keep = buf_ptr(out)
        check(lib().bz_ntt_result(self._h, int(buf_num), p, n))
        return out

    # ---- B200 additions
    def phase_times(self):
        ms, passes = ctypes.c_float(), ctypes.c_uint32()
        check(lib().bz_ntt_phase_times(self._h, ctypes.byref(ms), ctypes.byref(passes)))
        return {"total": ms.value, "passes": passes.value}

    def slot_device_ptr(self, buf_num):
        v = ctypes.c_uint64()
        check(lib().bz_ntt_slot_device_ptr(self._h, int(buf_num), ctypes.byref(v)))
        return v.value


class DistributedNTT:
    """Multi-GPU four-step NTT, one instance per rank (B200 addition, include/blaze_b200.h `bz_ntt_dist_*`).

    `exchange(handle_bytes) -> list of all ranks' handle bytes` is supplied by the caller (e.g.
    torch.distributed.all_gather_object); `barrier()` likewise.  With world == 1 both may be None."""

    def __init__(self, dclient: DriverClient, log_size, rank=0, world=1, field=2, inverse=False, exchange=None,
                 barrier=None):
        self.dclient = dclient
        self.rank, self.world, self.log_size = rank, world, log_size
        self.nbytes = (1 << log_size) * 32
        self._barrier = barrier or (lambda: None)
        h = ctypes.c_void_p()
        check(lib().bz_ntt_dist_new(dclient._h, int(field), int(log_size), 1 if inverse else 0, rank, world,
                                    ctypes.byref(h)))
        self._h = h
        # a ranked DriverClient (comm_init) exchanges the handles itself and supplies device-side barriers
        self._comm = world > 1 and dclient.comm_info() == (rank, world)
        if world > 1 and not self._comm:
            mine = ctypes.create_string_buffer(64)
            check(lib().bz_ntt_dist_ipc_handle(self._h, mine))
            handles = exchange(mine.raw)
            assert len(handles) == world
            blob = b"".join(handles)
            check(lib().bz_ntt_dist_open_peers(self._h, ctypes.c_char_p(blob)))
        self._barrier()

    def close(self):
        if getattr(self, "_h", None):
            self._barrier()
            lib().bz_ntt_dist_free(self._h)
            self._h = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().bz_ntt_dist_free(self._h)
                self._h = None
        except Exception:
            pass

    def set_input(self, full_input):
        p, n, keep = buf_ptr(full_input)
        check(lib().bz_ntt_dist_set_input(self._h, p, n))

    def get_output(self, full_output):
        p, n, keep = buf_ptr(full_output)
        check(lib().bz_ntt_dist_get_output(self._h, p, n))

    def run(self):
        """step1 (columns + twiddle + peer stores) -> barrier -> step3 (rows)."""
        if self._comm or self.world == 1:
            check(lib().bz_ntt_dist_run(self._h))      # stream-ordered, barriers on the device
            check(lib().bz_ntt_dist_sync(self._h))
            return
        self._barrier()                 # every rank's exchange buffer is free again
        check(lib().bz_ntt_dist_step1(self._h))
        check(lib().bz_ntt_dist_sync(self._h))
        self._barrier()                 # every rank has finished writing every exchange buffer
        check(lib().bz_ntt_dist_step3(self._h))
        check(lib().bz_ntt_dist_sync(self._h))

    def times(self):
        out = (ctypes.c_float * 2)()
        check(lib().bz_ntt_dist_times(self._h, out))
        return {"step1_ms": out[0], "step3_ms": out[1]}

    def plan(self):
        out = (ctypes.c_int32 * 4)()
        check(lib().bz_ntt_dist_plan(self._h, out))
        return {"log_n1": out[0], "log_n2": out[1], "column_passes": out[2], "row_passes": out[3]}

    def buffers(self):
        a, o, n = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        check(lib().bz_ntt_dist_buffers(self._h, ctypes.byref(a), ctypes.byref(o), ctypes.byref(n)))
        return a.value, o.value, n.value
