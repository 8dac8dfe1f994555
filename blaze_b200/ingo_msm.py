"""`MSMClient` and friends -- Python mirror of /root/reference/src/ingo_msm/{msm_api.rs,msm_cfg.rs}.

    MSMClient.new(MSMInit{mem_type, is_precompute, curve}, dclient)       msm_api.rs:44-55
    initialize(MSMParams) / start_process / set_data(MSMInput) / wait_result / result
"""
import ctypes
import enum
from dataclasses import dataclass
from typing import Optional, Tuple

from ._lib import lib, buf_ptr
from .driver_client import DriverClient, DriverPrimitive
from .error import check

PRECOMPUTE_FACTOR_BASE = 1     # msm_api.rs:39
PRECOMPUTE_FACTOR = 8          # msm_api.rs:40


class Curve(enum.IntEnum):     # msm_cfg.rs:4-8; numeric codes per msm_api.rs:359-364
    BLS377 = 0
    BN254 = 1
    BLS381 = 2


class PointMemoryType(enum.IntEnum):   # msm_cfg.rs:10-14
    HBM = 0
    DMA = 1


@dataclass
class MSMInit:                 # msm_api.rs:16-20
    mem_type: PointMemoryType
    is_precompute: bool
    curve: Curve


@dataclass
class MSMParams:               # msm_api.rs:22-26
    nof_elements: int
    hbm_point_addr: Optional[Tuple[int, int]] = None


@dataclass
class MSMInput:                # msm_api.rs:28-32
    points: Optional[object]
    scalars: object
    params: MSMParams


@dataclass
class MSMResult:               # msm_api.rs:33-37
    result: bytes
    result_label: int


def _hbm(params):
    if params.hbm_point_addr is None:
        return 0, 0, 0
    return 1, int(params.hbm_point_addr[0]), int(params.hbm_point_addr[1])


class MSMClient(DriverPrimitive):
    def __init__(self, init: MSMInit, dclient: DriverClient):
        self.driver_client = dclient
        self.mem_type = PointMemoryType(init.mem_type)
        self.precompute_factor = PRECOMPUTE_FACTOR if init.is_precompute else PRECOMPUTE_FACTOR_BASE
        self.curve = Curve(init.curve)
        h = ctypes.c_void_p()
        check(lib().bz_msm_new(dclient._h, int(self.curve), int(self.mem_type), 1 if init.is_precompute else 0,
                               ctypes.byref(h)))
        self._h = h
        s = [ctypes.c_uint32() for _ in range(4)]
        check(lib().bz_msm_sizes(self._h, *[ctypes.byref(x) for x in s]))
        self.scalar_size, self.point_size, self.result_point_size, _ = [x.value for x in s]

    @classmethod
    def new(cls, init, dclient):
        return cls(init, dclient)

    def close(self):
        if getattr(self, "_h", None):
            lib().bz_msm_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- DriverPrimitive
    def loaded_binary_parameters(self):                       # msm_api.rs:57-70
        out = (ctypes.c_uint32 * 2)()
        check(lib().bz_msm_loaded_binary_parameters(self._h, out))
        return [out[0], out[1]]

    def initialize(self, params: MSMParams):                  # msm_api.rs:72-111
        has, a, o = _hbm(params)
        check(lib().bz_msm_initialize(self._h, params.nof_elements, has, a, o))

    def start_process(self, param=None):                      # msm_api.rs:113-120
        check(lib().bz_msm_start_process(self._h))

    def set_data(self, data: MSMInput):                       # msm_api.rs:155-220
        pp, pn, k1 = buf_ptr(data.points)
        sp, sn, k2 = buf_ptr(data.scalars)
        has, a, o = _hbm(data.params)
        check(lib().bz_msm_set_data(self._h, pp, pn, sp, sn, data.params.nof_elements, has, a, o))

    def wait_result(self):                                    # msm_api.rs:222-238
        check(lib().bz_msm_wait_result(self._h))

    def result(self, param=None):                             # msm_api.rs:240-274
        out = bytearray(self.result_point_size)
        p, n, keep = buf_ptr(out)
        label = ctypes.c_uint32()
        check(lib().bz_msm_result(self._h, p, n, ctypes.byref(label)))
        return MSMResult(bytes(out), label.value)

    # ---- inherent methods
    def task_label(self):                                     # msm_api.rs:278-283
        v = ctypes.c_uint32()
        check(lib().bz_msm_task_label(self._h, ctypes.byref(v)))
        return v.value

    def nof_elements(self):                                   # msm_api.rs:285-290
        v = ctypes.c_uint32()
        check(lib().bz_msm_nof_elements(self._h, ctypes.byref(v)))
        return v.value

    def is_msm_engine_ready(self):                            # msm_api.rs:292-297
        v = ctypes.c_uint32()
        check(lib().bz_msm_is_msm_engine_ready(self._h, ctypes.byref(v)))
        return v.value

    def load_data_to_hbm(self, points, addr, offset):         # msm_api.rs:299-313
        p, n, keep = buf_ptr(points)
        check(lib().bz_msm_load_data_to_hbm(self._h, p, n, int(addr), int(offset)))

    def get_data_from_hbm(self, data_len, addr, offset):      # msm_api.rs:315-322
        out = bytearray(data_len)
        p, n, keep = buf_ptr(out)
        check(lib().bz_msm_get_data_from_hbm(self._h, p, n, int(addr), int(offset)))
        return bytes(out)

    def get_api(self):                                        # msm_api.rs:324-330: read every INGO_MSM_ADDR register
        """{register name: value} for every offset of msm_hw_code.rs:6-55 (the RESULT window 0x38..0xc4 is returned
        as bytes).  The reference reads and discards them; log_api_values() prints them."""
        regs = (ctypes.c_uint32 * 82)()
        check(lib().bz_msm_get_api(self._h, regs, 82))
        out = {}
        for name, off in INGO_MSM_ADDR.items():
            if name == "ADDR_HIF2CPU_C_RESULT":
                out[name] = bytes(bytearray(regs)[off:off + self.result_point_size])
            else:
                out[name] = regs[off // 4]
        return out

    def log_api_values(self):
        import logging
        for k, v in self.get_api().items():
            logging.getLogger("ingo_blaze").debug("%s: %s", k, v.hex() if isinstance(v, bytes) else hex(v))

    def table_build_ms(self):
        v = ctypes.c_float()
        check(lib().bz_msm_table_build_ms(self._h, ctypes.byref(v)))
        return v.value

    # ---- B200 additions
    def phase_times(self):
        out = (ctypes.c_float * 4)()
        check(lib().bz_msm_phase_times(self._h, out))
        return {"total": out[0], "sort": out[1], "accumulate": out[2], "reduce": out[3]}

    def set_window_bits(self, c):
        check(lib().bz_msm_set_window_bits(self._h, int(c)))

    def plan_info(self):
        out = (ctypes.c_uint32 * 8)()
        check(lib().bz_msm_plan_info_ex(self._h, out))
        return {"c": out[0], "windows": out[1], "buckets_per_window": out[2], "segment": out[3],
                "bucket_sets": out[4], "merged_table": bool(out[5]), "merged_table_mib": out[6],
                "accumulate": {0: "xyzz", 1: "batched-affine (multi-kernel)", 2: "batched-affine"}[(out[7] >> 24) & 0xF],
                "ba_rounds": (out[7] >> 28) & 0xF}

    def set_accumulate_mode(self, mode=-1, rounds=-1):
        """-1 automatic, 0 XYZZ mixed-add sweep, 2 fused batched-affine sweep (`rounds` tree rounds, -1 automatic)."""
        check(lib().bz_msm_set_accumulate_mode(self._h, int(mode), int(rounds)))

    def set_precompute(self, mode):
        """Window-merged table for HBM-resident points: 0 never, 1 from the second MSM on the same points
        (default), 2 immediately.  Same results in every mode."""
        check(lib().bz_msm_set_precompute(self._h, int(mode)))

    def set_raw_result(self, raw=True):
        """Leave result records projective (Z != 1, the reference's own format): for shards summed by combine_results."""
        check(lib().bz_msm_set_raw_result(self._h, 1 if raw else 0))

    def set_scalars_device(self, dev_ptr, params: MSMParams):
        has, a, o = _hbm(params)
        check(lib().bz_msm_set_scalars_device(self._h, int(dev_ptr), params.nof_elements, has, a, o))

    def combine_results(self, records: bytes, n: int) -> bytes:
        out = bytearray(self.result_point_size)
        p, ln, keep = buf_ptr(out)
        rp, rn, k2 = buf_ptr(records)
        check(lib().bz_msm_combine_results(self._h, rp, n, p, ln))
        return bytes(out)

    def generate_chain_points(self, p0q: bytes, first: int, n: int, addr: int, offset: int = 0):
        p, ln, keep = buf_ptr(p0q)
        check(lib().bz_msm_generate_chain_points(self._h, p, ln, int(first), int(n), int(addr), int(offset)))

    def expand_precompute(self, src_addr: int, n: int, dst_addr: int):
        """n factor-1 bases at src_addr -> the reference's x8 records (P, 2^32 P, ..) at dst_addr, on the device."""
        check(lib().bz_msm_expand_precompute(self._h, int(src_addr), int(n), int(dst_addr)))

    def field_selftest(self, a: bytes, b: bytes, n: int, op: int) -> bytes:
        out = bytearray(len(a))
        ap, _, k1 = buf_ptr(a)
        bp, _, k2 = buf_ptr(b)
        op_, _, k3 = buf_ptr(out)
        check(lib().bz_msm_field_selftest(self._h, ap, bp, op_, n, op))
        return bytes(out)


# msm_hw_code.rs:6-55
INGO_MSM_ADDR = {
    "ADDR_HIF2CPU_C_IMAGE_ID": 0x0, "ADDR_HIF2CPU_C_IMAGE_PARAMTERS": 0x4, "ADDR_HIF2CPU_C_MSM_ENGINE_READY": 0x8,
    "ADDR_HIF2CPU_C_MSM_TASK_LABEL": 0xc, "ADDR_CPU2HIF_C_BASES_HBM_START_ADDRESS_LO": 0x10,
    "ADDR_CPU2HIF_C_BASES_HBM_START_ADDRESS_HI": 0x14, "ADDR_CPU2HIF_C_BASES_SOURCE": 0x18,
    "ADDR_CPU2HIF_C_COEFFICIENTS_HBM_START_ADDRESS_LO": 0x1c, "ADDR_CPU2HIF_C_COEFFICIENTS_HBM_START_ADDRESS_HI": 0x20,
    "ADDR_CPU2HIF_C_COEFFICIENTS_SOURCE": 0x24, "ADDR_CPU2HIF_C_NUMBER_OF_MSM_ELEMENTS": 0x28,
    "ADDR_CPU2HIF_E_PUSH_MSM_TASK_TO_QUEUE": 0x2c, "ADDR_HIF2CPU_C_RESULT_VALID": 0x30, "ADDR_HIF2CPU_C_RESULT_LABEL": 0x34,
    "ADDR_HIF2CPU_C_RESULT": 0x38, "ADDR_CPU2HIF_E_POP_RESULT": 0xc8, "ADDR_HIF2CPU_C_NOF_PENDING_TASKS_IN_QUEUE": 0xcc,
    "ADDR_HIF2CPU_C_NOF_PENDING_RESULTS_IN_QUEUE": 0xd0, "ADDR_CPU2HIF_C_OPTIMIZATIONS": 0xd4,
    "ADDR_HIF2CPU_C_TASK_IN_FINAL_ACCUMULATION_PHASE": 0xd8, "ADDR_HIF2CPU_C_NOF_ELEMENTS_LEFT_IN_CURRENT_TASK": 0xdc,
    "ADDR_AXI2CPU_C_NUMBER_OF_COEFFICIENTS_IN_AXI_DMA_FIFO": 0xe0, "ADDR_AXI2CPU_C_NUMBER_OF_BASES_IN_AXI_DMA_FIFO": 0xe4,
    "ADDR_AXI2CPU_C_NUMBER_OF_COEFFICIENTS_IN_AXI_HBM_FIFO": 0xe8, "ADDR_AXI2CPU_C_NUMBER_OF_BASES_IN_AXI_HBM_FIFO": 0xec,
    "ADDR_HIF2CPU_E_BUCKET_ACCUMULATION_PHASE_COMPLETED": 0xf0, "ADDR_HIF2CPU_E_FINAL_ACCUMULATION_PHASE_COMPLETED": 0xf4,
    "ADDR_HIF2CPU_C_LAST_TASK_PHASE1_TOTAL_CLOCKS_LO": 0xf8, "ADDR_HIF2CPU_C_LAST_TASK_PHASE1_TOTAL_CLOCKS_HI": 0xfc,
    "ADDR_HIF2CPU_C_LAST_TASK_PHASE1_BUSY_ECADDER_CLOCKS_LO": 0x100, "ADDR_HIF2CPU_C_LAST_TASK_PHASE1_BUSY_ECADDER_CLOCKS_HI": 0x104,
    "ADDR_HIF2CPU_C_LAST_TASK_PHASE2_TOTAL_CLOCKS_LO": 0x108, "ADDR_HIF2CPU_C_LAST_TASK_PHASE2_TOTAL_CLOCKS_HI": 0x10c,
    "ADDR_HIF2CPU_C_LAST_TASK_PHASE2_BUSY_ECADDER_CLOCKS_LO": 0x110, "ADDR_HIF2CPU_C_LAST_TASK_PHASE2_BUSY_ECADDER_CLOCKS_HI": 0x114,
    "ADDR_HIF2CPU_C_LAST_TASK_PHASE3_TOTAL_CLOCKS_LO": 0x118, "ADDR_HIF2CPU_C_LAST_TASK_PHASE3_TOTAL_CLOCKS_HI": 0x11c,
    "ADDR_HIF2CPU_C_LAST_TASK_PHASE3_BUSY_ECADDER_CLOCKS_LO": 0x120, "ADDR_HIF2CPU_C_LAST_TASK_PHASE3_BUSY_ECADDER_CLOCKS_HI": 0x124,
    "ADDR_HIF2CPU_C_LAST_TASK_PHASE1_COEFFICIENTS_FIFO_BUSY_CLOCKS_LO": 0x128,
    "ADDR_HIF2CPU_C_LAST_TASK_PHASE1_COEFFICIENTS_FIFO_BUSY_CLOCKS_HI": 0x12c,
    "ADDR_HIF2CPU_C_LAST_TASK_PHASE1_COEFFICIENTS_FIFO_NOF_EMPTY_LO": 0x130,
    "ADDR_HIF2CPU_C_LAST_TASK_PHASE1_COEFFICIENTS_FIFO_NOF_EMPTY_HI": 0x134,
    "ADDR_HIF2CPU_C_LAST_TASK_PHASE1_BASES_FIFO_BUSY_CLOCKS_LO": 0x138, "ADDR_HIF2CPU_C_LAST_TASK_PHASE1_BASES_FIFO_BUSY_CLOCKS_HI": 0x13c,
    "ADDR_HIF2CPU_C_LAST_TASK_PHASE1_BASES_FIFO_NOF_EMPTY_LO": 0x140, "ADDR_HIF2CPU_C_LAST_TASK_PHASE1_BASES_FIFO_NOF_EMPTY_HI": 0x144,
}


class MSMImageParametrs:       # msm_api.rs:333-364 (sic: the reference's spelling)
    def __init__(self, is_stub, curve, number_of_ec_adders, buckets_mem_addr_width, number_of_segments, place_holder):
        self.hif2cpu_c_is_stub = is_stub
        self.hif2_cpu_c_curve = curve
        self.hif2_cpu_c_number_of_ec_adders = number_of_ec_adders
        self.hif2_cpu_c_buckets_mem_addr_width = buckets_mem_addr_width
        self.hif2_cpu_c_number_of_segments = number_of_segments
        self.hif2_cpu_c_place_holder = place_holder

    @staticmethod
    def parse_image_params(params: int):
        """`params.reverse_bits().to_be_bytes()` unpacked msb0 == LSB-first fields of `params`:
        bits 0..3 placeholder, 4..7 segments, 8..15 addr width, 16..19 ec adders, 20..27 curve,
        28..31 is_stub (msm_api.rs:336-354)."""
        return MSMImageParametrs(
            is_stub=(params >> 28) & 0xF, curve=(params >> 20) & 0xFF, number_of_ec_adders=(params >> 16) & 0xF,
            buckets_mem_addr_width=(params >> 8) & 0xFF, number_of_segments=(params >> 4) & 0xF,
            place_holder=params & 0xF)
