"""`PoseidonClient` -- Python mirror of /root/reference/src/ingo_hash/{poseidon_api.rs,utils.rs}."""
import ctypes
import enum
from dataclasses import dataclass

from ._lib import lib, buf_ptr
from .driver_client import DriverClient, DriverPrimitive
from .error import check


def num_of_elements_oct_tree(tree_height):          # utils.rs:2-10
    return sum(8 ** (tree_height - i - 1) for i in range(tree_height))


def num_of_elements_in_base_layer(tree_height):     # utils.rs:12-14
    return 8 ** (tree_height - 1)


class TreeMode(enum.IntEnum):                       # utils.rs:16-30
    TreeC = 0
    TreeD = 1

    @staticmethod
    def value(tree_mode):
        return int(tree_mode)


class Hash(enum.IntEnum):                           # poseidon_api.rs:11-13
    Poseidon = 0


@dataclass
class PoseidonInitializeParameters:                 # poseidon_api.rs:19-24
    tree_height: int
    tree_mode: TreeMode
    instruction_path: str = ""


@dataclass
class PoseidonResult:                               # poseidon_api.rs:26-30
    hash_byte: bytes
    hash_id: int
    layer_id: int

    @staticmethod
    def parse_poseidon_hash_results(data: bytes):   # poseidon_api.rs:42-71
        out = []
        assert len(data) % 64 == 0
        for i in range(0, len(data), 64):
            el = data[i:i + 64]
            meta = el[32:]
            hash_id = int.from_bytes(meta[:4], "little") & 0x3fffffff
            layer_id = int.from_bytes(meta[3:5] + b"\0\0", "little") >> 6
            out.append(PoseidonResult(bytes(el[:32]), hash_id, layer_id))
        return out


class PoseidonClient(DriverPrimitive):
    def __init__(self, ptype=Hash.Poseidon, dclient: DriverClient = None):
        self.dclient = dclient                      # pub field `dclient` (poseidon_api.rs:15-17)
        h = ctypes.c_void_p()
        check(lib().bz_poseidon_new(dclient._h, int(ptype), ctypes.byref(h)))
        self._h = h

    @classmethod
    def new(cls, ptype, dclient):
        return cls(ptype, dclient)

    def close(self):
        if getattr(self, "_h", None):
            lib().bz_poseidon_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def loaded_binary_parameters(self):                       # poseidon_api.rs:81-94
        out = (ctypes.c_uint32 * 2)()
        check(lib().bz_poseidon_loaded_binary_parameters(self._h, out))
        return [out[0], out[1]]

    def initialize(self, param: PoseidonInitializeParameters):   # poseidon_api.rs:96-111
        check(lib().bz_poseidon_initialize(self._h, param.tree_height, int(param.tree_mode),
                                           (param.instruction_path or "").encode()))

    def set_data(self, input):                                # poseidon_api.rs:117-122
        p, n, keep = buf_ptr(input)
        check(lib().bz_poseidon_set_data(self._h, p, n))

    def start_process(self, param=None):                      # todo!() in the reference
        check(lib().bz_poseidon_start_process(self._h))

    def wait_result(self):                                    # todo!() in the reference
        check(lib().bz_poseidon_wait_result(self._h))

    def result(self, expected_result=None):                   # poseidon_api.rs:128-146
        if expected_result is None:
            from .error import InvalidPrimitiveParam
            raise InvalidPrimitiveParam("result needs Some(expected_result)", -4)
        cap = max(int(expected_result), self.get_num_of_pending_results()) + 8
        out = bytearray(cap * 64)
        p, n, keep = buf_ptr(out)
        got = ctypes.c_size_t()
        check(lib().bz_poseidon_result(self._h, int(expected_result), p, cap, ctypes.byref(got)))
        return PoseidonResult.parse_poseidon_hash_results(bytes(out[:got.value * 64]))

    def get_last_element_sent_to_ring(self):                  # poseidon_api.rs:149-154
        v = ctypes.c_uint32()
        check(lib().bz_poseidon_get_last_element_sent_to_ring(self._h, ctypes.byref(v)))
        return v.value

    def get_num_of_pending_results(self):                     # poseidon_api.rs:156-161
        v = ctypes.c_uint32()
        check(lib().bz_poseidon_get_num_of_pending_results(self._h, ctypes.byref(v)))
        return v.value

    def get_raw_results(self, num_of_results):                # poseidon_api.rs:191-196
        out = bytearray(64 * num_of_results)
        p, n, keep = buf_ptr(out)
        check(lib().bz_poseidon_get_raw_results(self._h, int(num_of_results), p))
        return bytes(out)

    def get_last_hash_sent_to_host(self):                     # poseidon_api.rs:198-203
        v = ctypes.c_uint32()
        check(lib().bz_poseidon_get_last_hash_sent_to_host(self._h, ctypes.byref(v)))
        return v.value

    # ---- B200 additions
    def device_ms(self):
        """Kernel milliseconds since initialize() (the analogue of the core's clock counters)."""
        v = ctypes.c_float()
        check(lib().bz_poseidon_device_ms(self._h, ctypes.byref(v)))
        return v.value

    def permute(self, states: bytes, t: int, mds_mode: int = 0) -> bytes:
        """Bare permutation of len(states) / (32 t) states of width t (3, 9, 12); mds_mode 1 = Grain-sampled MDS."""
        n = len(states) // (32 * t)
        out = bytearray(n * t * 32)
        ip, _, k1 = buf_ptr(states)
        op, _, k2 = buf_ptr(out)
        check(lib().bz_poseidon_permute(self._h, t, mds_mode, ip, n, op))
        return bytes(out)

    def log_api_values(self):                                 # poseidon_api.rs:245-253
        return {"last_element": self.get_last_element_sent_to_ring(), "last_hash": self.get_last_hash_sent_to_host()}
