"""ctypes wrapper of oracle/_build/libplanexec.so - TEST INFRASTRUCTURE ONLY.

The CPU interpreter of the device layer plan (oracle/plan_exec.cc): same data structures as the CUDA
engine, plain loops.  Used by tests/ to validate nhans_b200/csrc/plan.cc against the oracle without a GPU
and to isolate kernel bugs from plan bugs on the GPU box."""
from __future__ import annotations

import ctypes
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libplanexec.so")


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def _load():
    if not os.path.exists(_SO):
        build()
    lib = ctypes.CDLL(_SO)
    lib.planexec_create.restype = ctypes.c_void_p
    lib.planexec_error.restype = ctypes.c_char_p
    lib.planexec_plan_json.restype = ctypes.c_char_p
    lib.planexec_plan_json.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.planexec_destroy.argtypes = [ctypes.c_void_p]
    return lib


def marshal_weights(weights):
    names = sorted(weights)
    arrs = [np.ascontiguousarray(weights[n], dtype=np.float32) for n in names]
    c_names = (ctypes.c_char_p * len(names))(*[n.encode() for n in names])
    c_sizes = (ctypes.c_int64 * len(names))(*[a.size for a in arrs])
    c_data = (ctypes.c_void_p * len(names))(*[a.ctypes.data for a in arrs])
    return names, arrs, c_names, c_sizes, c_data


class PlanExec:
    def __init__(self, weights, variant=0, win_cap=4, row_cap=1):
        self.lib = _load()
        names, arrs, c_names, c_sizes, c_data = marshal_weights(weights)
        self.h = self.lib.planexec_create(c_names, c_sizes, c_data, len(names), variant, win_cap, row_cap)
        if not self.h:
            raise RuntimeError(self.lib.planexec_error().decode())

    def close(self):
        if self.h:
            self.lib.planexec_destroy(self.h)
            self.h = None

    def plan(self, net=0):
        return json.loads(self.lib.planexec_plan_json(self.h, net).decode())

    def embed(self, ctx_logmag):
        x = np.ascontiguousarray(ctx_logmag, np.float32)
        R = x.shape[0]
        out = np.zeros((R, 512), np.float32)
        self.lib.planexec_embed(ctypes.c_void_p(self.h), x.ctypes.data_as(ctypes.c_void_p), R, out.ctypes.data_as(ctypes.c_void_p))
        return out

    def masknet(self, logmag, frame_offs, emb_a, emb_b):
        lm = np.ascontiguousarray(logmag, np.float32)
        fo = np.ascontiguousarray(frame_offs, np.int64)
        ea = np.ascontiguousarray(emb_a, np.float32)
        eb = np.ascontiguousarray(emb_b, np.float32)
        out = np.zeros_like(lm)
        self.lib.planexec_masknet(ctypes.c_void_p(self.h), lm.ctypes.data_as(ctypes.c_void_p), fo.ctypes.data_as(ctypes.c_void_p),
                                  len(fo) - 1, ea.ctypes.data_as(ctypes.c_void_p), eb.ctypes.data_as(ctypes.c_void_p),
                                  out.ctypes.data_as(ctypes.c_void_p))
        return out

    def read_buffer(self, net, buf):
        g = self.plan(net)["bufs"][buf]
        n = g["pixels"] * g["C"]
        out = np.zeros(n, np.float32)
        rc = self.lib.planexec_read_buffer(ctypes.c_void_p(self.h), net, buf, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(n))
        if rc:
            raise RuntimeError("no such buffer")
        return out.reshape(g["pixels"], g["C"])


def grid_gather(grid, flat, n_units):
    """Logical [n, H, W, C] view of a padded grid buffer `flat` [pixels, C] (plan.h grid_pixel)."""
    H, W, C = grid["H"], grid["W"], grid["C"]
    n = np.arange(n_units)[:, None, None]
    h = np.arange(H)[None, :, None]
    w = np.arange(W)[None, None, :]
    if grid["mode"] == 1:
        pix = (n * W + w) * H + h
    else:
        y = h + grid["oy"]
        x = w + grid["ox"]
        plane = (y % grid["sh"]) * grid["sw"] + (x % grid["sw"])
        pix = plane * grid["plane_stride"] + n * grid["ustride"] + (y // grid["sh"]) * grid["rstride"] + (x // grid["sw"])
    return flat[pix]
