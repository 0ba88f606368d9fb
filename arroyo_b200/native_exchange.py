"""ctypes front of the opt-in native Shuffle edge (csrc/exchange.cu, libarroyo_b200_xchg.so): the round of
multi_gpu.ShuffleExchange.round_packed as ONE C call (partition + control all-gather + variable all-to-all, NCCL from
C++).  Not measured yet -- written at the end of round 1 as the first experiment of round 2."""
import ctypes as C
import os
from typing import Optional

from . import ffi
from .context import WatermarkHolder

NO_WM = -(1 << 63)
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    from . import build_exchange
    path = build_exchange.LIB
    if not os.path.exists(path):
        raise ffi.ArroyoB200Error(ffi.FATAL, "libarroyo_b200_xchg.so has not been built (arroyo_b200.build_exchange)")
    ffi.load()  # the exchange library links against libarroyo_b200.so
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    lib.arroyo_b200_xchg_unique_id.restype = C.c_int32
    lib.arroyo_b200_xchg_unique_id.argtypes = [C.c_void_p]
    lib.arroyo_b200_xchg_create.restype = C.c_int32
    lib.arroyo_b200_xchg_create.argtypes = [C.c_int32, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                            C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]
    lib.arroyo_b200_xchg_destroy.argtypes = [C.c_void_p]
    lib.arroyo_b200_xchg_last_error.restype = C.c_char_p
    lib.arroyo_b200_xchg_last_error.argtypes = [C.c_void_p]
    lib.arroyo_b200_xchg_round.restype = C.c_int32
    lib.arroyo_b200_xchg_round.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int64, C.c_int64, C.c_int32,
                                           C.POINTER(C.c_uint64), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                           C.POINTER(C.c_int32)]
    _lib = lib
    return lib


class NativeExchange:
    """Same contract as ShuffleExchange.round_packed, fed with raw column pointers (the partition happens inside)."""

    def __init__(self, torch, dist, rank: int, world: int, device_index: int, stream: int, n_cols: int, key_col: int,
                 max_rows: int, max_recv_rows: int):
        self.lib = load()
        self.world, self.n_cols, self.rank = world, n_cols, rank
        # the NCCL unique id travels over the process group that already exists
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            st = self.lib.arroyo_b200_xchg_unique_id(buf)
            if st != ffi.OK:
                raise ffi.ArroyoB200Error(st, "ncclGetUniqueId failed")
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.to(torch.device("cuda", device_index))
        dist.broadcast(uid, src=0)
        raw = (C.c_uint8 * 128)(*uid.cpu().tolist())
        self.h = C.c_void_p()
        st = self.lib.arroyo_b200_xchg_create(device_index, stream, rank, world, raw, n_cols, key_col, max_rows,
                                              max_recv_rows, C.byref(self.h))
        if st != ffi.OK:
            raise ffi.ArroyoB200Error(st, "arroyo_b200_xchg_create failed")
        self.holder = WatermarkHolder(world)
        self.ctrl = None  # (ShuffleExchange compatibility: nothing to write counts into)
        self.bytes_sent = 0
        self._cols = (C.c_uint64 * (world * n_cols))()
        self._rows = (C.c_int64 * world)()
        self._wms = (C.c_int64 * world)()
        # One throw-away round with a few rows for every destination, now, while the devices are quiet: NCCL sets its
        # peer-to-peer connections up at the first send / receive between two ranks, and that set-up (allocations, IPC
        # handles, a handshake per peer) should not run in the middle of the pipeline with every SM taken.  Nothing of
        # it reaches an operator; no watermark is attached.
        if world > 1:
            n = int(min(max_rows, 64 * world))
            dev = torch.device("cuda", device_index)
            self._warm = [torch.arange(n, dtype=torch.int64, device=dev) * 0x1E3779B97F4A7C15 + rank for _ in range(n_cols)]
            torch.cuda.synchronize(dev)
            self.round_packed([t.data_ptr() for t in self._warm], None, n, None)
            torch.cuda.synchronize(dev)
            self.bytes_sent = 0

    def round_packed(self, col_ptrs, _counts, n_rows: int, watermark: Optional[int], more: bool = False):
        inp = (C.c_uint64 * self.n_cols)(*col_ptrs) if n_rows > 0 else None
        any_more = C.c_int32(0)
        wm = NO_WM if watermark is None else int(min(watermark, (1 << 63) - 1))
        st = self.lib.arroyo_b200_xchg_round(self.h, inp, n_rows, wm, 1 if more else 0, self._cols, self._rows, self._wms,
                                             C.byref(any_more))
        if st != ffi.OK:
            raise ffi.ArroyoB200Error(st, (self.lib.arroyo_b200_xchg_last_error(self.h) or b"").decode())
        nc = self.n_cols
        self.bytes_sent += (n_rows - int(self._rows[self.rank])) * nc * 8  # rows that left for another rank
        batches = [([self._cols[s * nc + c] for c in range(nc)], int(self._rows[s])) for s in range(self.world) if self._rows[s]]
        before = self.holder.last_present_watermark
        for s in range(self.world):
            if self._wms[s] != NO_WM:
                self.holder.set(s, int(self._wms[s]))
        after = self.holder.last_present_watermark
        return batches, (after if after is not None and after != before else None), bool(any_more.value)

    def close(self):
        if self.h:
            self.lib.arroyo_b200_xchg_destroy(self.h)
            self.h = C.c_void_p()
