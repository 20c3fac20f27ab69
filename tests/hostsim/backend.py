"""Host-simulator backend for pof.sharded.ShardedPass (TEST INFRASTRUCTURE): runs the three shard stages with the
device code compiled for the CPU (libhostsim.so), so the multi-rank orchestration can be tested with gloo and no GPU."""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_P = ctypes.POINTER(ctypes.c_double)


def load():
    lib = ctypes.CDLL(os.path.join(_HERE, "libhostsim.so"))
    lib.hs_ws_create.restype = ctypes.c_void_p
    lib.hs_ws_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_void_p]
    lib.hs_ws_free.argtypes = [ctypes.c_void_p]
    lib.hs_stage_a.argtypes = [ctypes.c_void_p] * 4
    lib.hs_stage_b.argtypes = [ctypes.c_void_p] * 9
    lib.hs_stage_c.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_void_p,
                               ctypes.c_void_p, ctypes.c_void_p]
    lib.hs_filter_chain.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.hs_smooth_chain.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    # the same entry points on the large-state ("tile") family
    lib.hs_tile_ws_create.restype = ctypes.c_void_p
    lib.hs_tile_ws_create.argtypes = lib.hs_ws_create.argtypes
    lib.hs_tile_ws_free.argtypes = [ctypes.c_void_p]
    lib.hs_tile_stage_a.argtypes = lib.hs_stage_a.argtypes
    lib.hs_tile_stage_b.argtypes = lib.hs_stage_b.argtypes
    lib.hs_tile_stage_c.argtypes = lib.hs_stage_c.argtypes
    lib.hs_tile_filter_chain.argtypes = lib.hs_filter_chain.argtypes
    lib.hs_tile_smooth_chain.argtypes = lib.hs_smooth_chain.argtypes
    return lib


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class HostBackend:
    def __init__(self, d, q, n_loc, chunk_len, qL, family="thread"):
        lib = load()
        self.qL = np.ascontiguousarray(qL, dtype=np.float64)
        if family == "tile":  # CTA-per-chunk / CTA-per-node device code (csrc/pof_tile.cuh), any (d, q)
            class _Tile:
                hs_stage_a, hs_stage_b, hs_stage_c = lib.hs_tile_stage_a, lib.hs_tile_stage_b, lib.hs_tile_stage_c
                hs_filter_chain, hs_smooth_chain = lib.hs_tile_filter_chain, lib.hs_tile_smooth_chain
            self.lib = _Tile
            self.ws = lib.hs_tile_ws_create(d, q, n_loc, chunk_len, self.qL.ctypes.data_as(ctypes.c_void_p))
        else:
            self.lib = lib
            self.ws = lib.hs_ws_create(d, q, n_loc, chunk_len, self.qL.ctypes.data_as(ctypes.c_void_p))
        assert self.ws

    def stage_a(self, H, c, carry_f):
        assert self.lib.hs_stage_a(self.ws, _p(H), _p(c), _p(carry_f)) == 0

    def stage_b(self, H, c, state_in, fmeans, fchols, carry_s, state_end, partials):
        # the outputs are views into one payload tensor: write through contiguous temporaries
        cs, se, pa = carry_s.clone(), state_end.clone(), partials.clone()
        assert self.lib.hs_stage_b(self.ws, _p(H), _p(c), _p(state_in), _p(fmeans), _p(fchols), _p(cs), _p(se),
                                   _p(pa)) == 0
        carry_s.copy_(cs), state_end.copy_(se), partials.copy_(pa)

    def stage_c(self, seed, is_last, has_row0, cscale, means, chols, partials2):
        assert self.lib.hs_stage_c(self.ws, _p(seed), int(has_row0), float(cscale[0]), _p(means), _p(chols),
                                   _p(partials2)) == 0

    def filter_chain(self, D, count, state_in, elems, state_out):
        assert self.lib.hs_filter_chain(D, count, _p(state_in), _p(elems), _p(state_out)) == 0

    def smooth_chain(self, D, count, state_in, elems, state_out):
        assert self.lib.hs_smooth_chain(D, count, _p(state_in), _p(elems), _p(state_out)) == 0
