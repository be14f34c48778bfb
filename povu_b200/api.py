"""ctypes binding of the povu_b200 C ABI (include/povu_b200.h) and the host-side mirror of the
reference's operator interface for the decompose -> call path.

The product path loads ONLY the CUDA library (libpovu_b200.so, sm_100a).  There is no CPU fallback:
if the library or a CUDA device is missing, construction raises.  (tests/emu builds a host debug
rehearsal of the same kernel sources; it is loaded only by tests that pass `lib_path` explicitly.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libpovu_b200.so")

_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def _as_ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


def load_library(lib_path=None):
    path = lib_path or DEFAULT_LIB
    if not os.path.exists(path):
        raise RuntimeError(
            f"povu_b200: CUDA library {path} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback.")
    lib = C.CDLL(path)
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    P32, P8 = C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
    lib.pvgpu_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    lib.pvgpu_create.restype = C.c_int
    lib.pvgpu_destroy.argtypes = [vp]
    lib.pvgpu_last_error.argtypes = [vp]
    lib.pvgpu_last_error.restype = C.c_char_p
    lib.pvgpu_launch_count.argtypes = [vp]
    lib.pvgpu_launch_count.restype = u64
    lib.pvgpu_host_sync_count.argtypes = [vp]
    lib.pvgpu_host_sync_count.restype = u64
    lib.pvgpu_peak_device_bytes.argtypes = [vp]
    lib.pvgpu_peak_device_bytes.restype = u64
    lib.pvgpu_build_kind.restype = C.c_int
    lib.pvgpu_set_profiling.argtypes = [vp, C.c_int]
    lib.pvgpu_profile_json.argtypes = [vp, C.c_char_p, C.c_size_t]
    lib.pvgpu_profile_json.restype = C.c_size_t
    lib.pvgpu_selftest_prims.argtypes = [vp, u64, u64, C.c_int, C.POINTER(u64)]
    lib.pvgpu_selftest_prims.restype = C.c_int
    lib.pvgpu_selftest_arcs.argtypes = [vp, u32, u32, u64, C.POINTER(u64), C.POINTER(C.c_int)]
    lib.pvgpu_selftest_arcs.restype = C.c_int
    lib.pvgpu_selftest_text.argtypes = [vp, u32, C.POINTER(u64)]
    lib.pvgpu_decompose.argtypes = [vp, u32, u64, vp, vp, u32, C.POINTER(vp)]
    lib.pvgpu_decompose.restype = C.c_int
    lib.pvgpu_decompose_dev.argtypes = [vp, u32, u64, vp, vp, u32, C.POINTER(vp)]
    lib.pvgpu_decompose_dev.restype = C.c_int
    lib.pvgpu_flubbles_count.argtypes = [vp]
    lib.pvgpu_flubbles_count.restype = u64
    lib.pvgpu_flubbles_component_count.argtypes = [vp]
    lib.pvgpu_flubbles_component_count.restype = u32
    lib.pvgpu_flubbles_dfs_rounds.argtypes = [vp]
    lib.pvgpu_flubbles_dfs_rounds.restype = u32
    lib.pvgpu_flubbles_hash_collisions.argtypes = [vp]
    lib.pvgpu_flubbles_hash_collisions.restype = u32
    lib.pvgpu_flubbles_arc_rounds.argtypes = [vp]
    lib.pvgpu_flubbles_arc_rounds.restype = u32
    lib.pvgpu_flubbles_faithful_classes.argtypes = [vp]
    lib.pvgpu_flubbles_faithful_classes.restype = u32
    lib.pvgpu_flubbles_copy.argtypes = [vp, P32, P32, P8, P32, P8, P32, P32]
    lib.pvgpu_flubbles_copy_components.argtypes = [vp, P32, P32, P32, P32, P32]
    lib.pvgpu_flubbles_stack_size.argtypes = [vp]
    lib.pvgpu_flubbles_stack_size.restype = u64
    lib.pvgpu_flubbles_copy_stack.argtypes = [vp, P32, P32, P32]
    lib.pvgpu_flubbles_free.argtypes = [vp]
    sz = C.c_size_t
    lib.pvgpu_session_open.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]
    lib.pvgpu_session_open.restype = C.c_int
    lib.pvgpu_session_close.argtypes = [vp]
    lib.pvgpu_session_open_hostparse.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]
    lib.pvgpu_session_open_hostparse.restype = C.c_int
    lib.pvgpu_session_gfa_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(C.c_double)]
    lib.pvgpu_session_copy_graph.argtypes = [vp] + [vp] * 9
    lib.pvgpu_session_copy_graph.restype = C.c_int
    lib.pvgpu_session_label_bytes.argtypes = [vp]
    lib.pvgpu_session_label_bytes.restype = u64
    lib.pvgpu_session_path_name.argtypes = [vp, u32]
    lib.pvgpu_session_path_name.restype = C.c_char_p
    lib.pvgpu_session_from_arrays.argtypes = [vp, u32, vp, vp, vp, vp, u64, vp, vp, u32, C.POINTER(C.c_char_p), vp, vp, C.POINTER(vp)]
    lib.pvgpu_session_from_arrays.restype = C.c_int
    lib.pvgpu_session_hot_path.argtypes = [vp, C.POINTER(C.c_char_p), sz, C.c_int, C.POINTER(C.c_double)]
    lib.pvgpu_session_hot_path.restype = C.c_int
    lib.pvgpu_session_sizes.argtypes = [vp] + [C.POINTER(u64)] * 4
    lib.pvgpu_session_decompose.argtypes = [vp]
    lib.pvgpu_session_decompose.restype = C.c_int
    lib.pvgpu_session_flubble_count.argtypes = [vp]
    lib.pvgpu_session_flubble_count.restype = u64
    lib.pvgpu_session_tree_count.argtypes = [vp]
    lib.pvgpu_session_tree_count.restype = u32
    lib.pvgpu_session_tree_component.argtypes = [vp, u32]
    lib.pvgpu_session_tree_component.restype = u32
    lib.pvgpu_session_pvst_text.argtypes = [vp, u32, C.POINTER(sz)]
    lib.pvgpu_session_pvst_text.restype = vp
    lib.pvgpu_session_write_pvst.argtypes = [vp, C.c_char_p]
    lib.pvgpu_session_write_pvst.restype = C.c_int
    lib.pvgpu_session_read_pvst.argtypes = [vp, C.c_char_p]
    lib.pvgpu_session_read_pvst.restype = C.c_int
    lib.pvgpu_session_call.argtypes = [vp, C.POINTER(C.c_char_p), sz]
    lib.pvgpu_session_call.restype = C.c_int
    lib.pvgpu_session_set_restrict.argtypes = [vp, C.c_char_p]
    lib.pvgpu_session_set_restrict.restype = C.c_int
    lib.pvgpu_session_record_count.argtypes = [vp]
    lib.pvgpu_session_record_count.restype = u64
    lib.pvgpu_session_call_stats.argtypes = [vp, C.POINTER(u64)]
    lib.pvgpu_session_vcf_text.argtypes = [vp, sz, C.c_int, C.c_char_p, C.POINTER(sz)]
    lib.pvgpu_session_vcf_text.restype = vp
    lib.pvgpu_session_vcf_write.argtypes = [vp, sz, C.c_int, C.c_char_p, C.c_int, C.POINTER(u64)]
    lib.pvgpu_session_vcf_write.restype = C.c_int
    lib.pvgpu_free.argtypes = [vp]
    lib.pvgpu_session_shard.argtypes = [vp, u32, u32, C.POINTER(u32)]
    lib.pvgpu_session_shard.restype = C.c_int
    lib.pvgpu_session_records_packed_bytes.argtypes = [vp]
    lib.pvgpu_session_records_packed_bytes.restype = u64
    lib.pvgpu_session_records_pack.argtypes = [vp, vp, u64, C.POINTER(u64)]
    lib.pvgpu_session_records_pack.restype = C.c_int
    lib.pvgpu_session_reshard.argtypes = [vp, u32, u32, C.c_int, C.POINTER(u32), C.POINTER(C.c_double)]
    lib.pvgpu_session_reshard.restype = C.c_int
    lib.pvgpu_session_tree_regions.argtypes = [vp, vp, vp]
    lib.pvgpu_session_record_lines.argtypes = [vp, C.POINTER(sz), C.POINTER(vp), C.POINTER(vp), C.POINTER(sz)]
    lib.pvgpu_session_record_lines.restype = vp
    return lib


def edge_ports(g):
    """(ep1, ep2) u32 port arrays of a synth/loader graph dict: port = 2*vertex + side."""
    ep1 = (g["v1"].astype(np.uint32) * 2 + g["s1"].astype(np.uint32)).astype(np.uint32)
    ep2 = (g["v2"].astype(np.uint32) * 2 + g["s2"].astype(np.uint32)).astype(np.uint32)
    return np.ascontiguousarray(ep1), np.ascontiguousarray(ep2)


class Flubbles:
    """Result of decompose: flat flubble table + per-component table (see include/povu_b200.h)."""

    def __init__(self, lib, h):
        n = lib.pvgpu_flubbles_count(h)
        nc = lib.pvgpu_flubbles_component_count(h)
        u32, u8 = np.uint32, np.uint8
        self.comp = np.zeros(n, u32); self.a_v = np.zeros(n, u32); self.z_v = np.zeros(n, u32)
        self.a_fwd = np.zeros(n, u8); self.z_fwd = np.zeros(n, u8)
        self.parent = np.zeros(n, u32); self.depth = np.zeros(n, u32)
        lib.pvgpu_flubbles_copy(h, _as_ptr(self.comp, C.c_uint32), _as_ptr(self.a_v, C.c_uint32), _as_ptr(self.a_fwd, C.c_uint8),
                                _as_ptr(self.z_v, C.c_uint32), _as_ptr(self.z_fwd, C.c_uint8), _as_ptr(self.parent, C.c_uint32),
                                _as_ptr(self.depth, C.c_uint32))
        self.comp_root = np.zeros(nc, u32); self.comp_nv = np.zeros(nc, u32); self.comp_nfl = np.zeros(nc, u32)
        self.comp_has_tips = np.zeros(nc, u32); self.comp_nonlaminar = np.zeros(nc, u32)
        lib.pvgpu_flubbles_copy_components(h, _as_ptr(self.comp_root, C.c_uint32), _as_ptr(self.comp_nv, C.c_uint32),
                                           _as_ptr(self.comp_nfl, C.c_uint32), _as_ptr(self.comp_has_tips, C.c_uint32),
                                           _as_ptr(self.comp_nonlaminar, C.c_uint32))
        ns = lib.pvgpu_flubbles_stack_size(h)
        self.stack_v = np.zeros(ns, u32); self.stack_class = np.zeros(ns, u32); self.stack_fwd = np.zeros(ns, u32)
        if ns:
            lib.pvgpu_flubbles_copy_stack(h, _as_ptr(self.stack_v, C.c_uint32), _as_ptr(self.stack_class, C.c_uint32),
                                          _as_ptr(self.stack_fwd, C.c_uint32))
        self.dfs_rounds = lib.pvgpu_flubbles_dfs_rounds(h)
        self.arc_rounds = lib.pvgpu_flubbles_arc_rounds(h)
        self.faithful_classes = lib.pvgpu_flubbles_faithful_classes(h)
        self.n = int(n)

    def labels(self, ids):
        """Flubble::as_str (pvst.hpp:408-412): e.g. '>12>17'."""
        a = ids[self.a_v]; z = ids[self.z_v]
        return [f"{'>' if af else '<'}{x}{'>' if zf else '<'}{y}" for x, af, y, zf in zip(a, self.a_fwd, z, self.z_fwd)]

    def pvst_texts(self, ids):
        """{component_id: text} of the .pvst v0.0.3 files the reference would write
        (to_pvst.cpp:23-109; components with < 3 vertices are skipped, decompose.cpp:135-142)."""
        out = {}
        labels = self.labels(ids)
        base = 0
        for k in range(len(self.comp_nv)):
            nfl = int(self.comp_nfl[k])
            if self.comp_nv[k] >= 3:
                ch = [[] for _ in range(nfl + 1)]
                par = self.parent[base:base + nfl]
                for j in range(nfl):
                    ch[par[j]].append(j + 1)
                lines = ["H\t0.0.3\t.\t.\t.\n"]
                for i in range(nfl + 1):
                    kids = ", ".join(map(str, ch[i])) if ch[i] else "."
                    if i == 0:
                        lines.append(f"D\t0\t.\t{kids}\t.\n")
                    else:
                        lines.append(f"F\t{i}\t{labels[base + i - 1]}\t{kids}\tL\n")
                out[k + 1] = "".join(lines)
            base += nfl
        return out


class Povu:
    """One context = one CUDA device + stream."""

    def __init__(self, device=0, stream=None, lib_path=None):
        self.lib = load_library(lib_path)
        kind = self.lib.pvgpu_build_kind()
        if lib_path is None and kind != 1:
            raise RuntimeError("povu_b200: the default library must be the CUDA build")
        h = C.c_void_p()
        rc = self.lib.pvgpu_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise RuntimeError("povu_b200: pvgpu_create failed (no usable CUDA device; there is no CPU fallback)")
        self.h = h

    def close(self):
        if self.h:
            self.lib.pvgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(self.lib.pvgpu_launch_count(self.h))

    @property
    def host_syncs(self):
        return int(self.lib.pvgpu_host_sync_count(self.h))

    @property
    def peak_bytes(self):
        return int(self.lib.pvgpu_peak_device_bytes(self.h))

    def selftest_text(self, max_an=700):
        bad = C.c_uint64()
        if self.lib.pvgpu_selftest_text(self.h, int(max_an), C.byref(bad)) != 0:
            raise RuntimeError(self._err())
        return int(bad.value)

    def selftest_arcs(self, n, n_classes, seed=1):
        """Fired-arc fixpoint vs the reference's stack machine on a random class sequence; returns (mismatches, rounds)."""
        bad, rounds = C.c_uint64(0), C.c_int(0)
        if self.lib.pvgpu_selftest_arcs(self.h, int(n), int(n_classes), int(seed), C.byref(bad), C.byref(rounds)) != 0:
            raise RuntimeError("povu_b200: " + self.lib.pvgpu_last_error(self.h).decode())
        return int(bad.value), int(rounds.value)

    def selftest_prims(self, n, seed=1, bits=32):
        """Device primitives (pair sort, scans, RMQ) against the host's std:: algorithms; returns the mismatch count."""
        bad = C.c_uint64(0)
        if self.lib.pvgpu_selftest_prims(self.h, int(n), int(seed), int(bits), C.byref(bad)) != 0:
            raise RuntimeError("povu_b200: " + self.lib.pvgpu_last_error(self.h).decode())
        return int(bad.value)

    def set_profiling(self, on):
        self.lib.pvgpu_set_profiling(self.h, 1 if on else 0)

    def profile(self):
        """Per-kernel CUDA-event timings accumulated since set_profiling(True)."""
        import json
        n = self.lib.pvgpu_profile_json(self.h, None, 0)
        buf = C.create_string_buffer(n + 16)
        self.lib.pvgpu_profile_json(self.h, buf, n + 16)
        return json.loads(buf.value.decode())

    def _err(self):
        return self.lib.pvgpu_last_error(self.h).decode()

    def decompose(self, n_vertices, ep1, ep2, want_stack=False):
        """Host arrays in, Flubbles out (H2D + kernels + D2H)."""
        ep1 = np.ascontiguousarray(ep1, dtype=np.uint32); ep2 = np.ascontiguousarray(ep2, dtype=np.uint32)
        out = C.c_void_p()
        rc = self.lib.pvgpu_decompose(self.h, int(n_vertices), len(ep1), ep1.ctypes.data, ep2.ctypes.data, 1 if want_stack else 0, C.byref(out))
        if rc != 0:
            raise RuntimeError("pvgpu_decompose: " + self._err())
        try:
            return Flubbles(self.lib, out)
        finally:
            self.lib.pvgpu_flubbles_free(out)

    def decompose_dev(self, n_vertices, n_edges, d_ep1, d_ep2, want_stack=False, fetch=True):
        """Device pointers in (ints)."""
        out = C.c_void_p()
        rc = self.lib.pvgpu_decompose_dev(self.h, int(n_vertices), int(n_edges), C.c_void_p(d_ep1), C.c_void_p(d_ep2), 1 if want_stack else 0, C.byref(out))
        if rc != 0:
            raise RuntimeError("pvgpu_decompose_dev: " + self._err())
        try:
            return Flubbles(self.lib, out) if fetch else int(self.lib.pvgpu_flubbles_count(out))
        finally:
            self.lib.pvgpu_flubbles_free(out)


class Session:
    """One GFA loaded on one device: decompose -> (.pvst) -> call -> VCF text.  Mirrors the reference's
    `povu decompose / call / gfa2vcf` drivers (app/subcommand/*.cpp) through the pvgpu_session_* C ABI."""

    def __init__(self, pv, gfa_path=None, inc_vtx_labels=True, inc_refs=True, _handle=None, host_parser=False):
        """The GFA text is parsed on the device (pv_gfa.cuh); host_parser=True is the test hook that loads it with the
        host loader instead (the checker of the device parser)."""
        self.pv, self.lib = pv, pv.lib
        if _handle is not None:
            self.h = _handle
            return
        h = C.c_void_p()
        opener = self.lib.pvgpu_session_open_hostparse if host_parser else self.lib.pvgpu_session_open
        rc = opener(pv.h, os.fsencode(gfa_path), int(inc_vtx_labels), int(inc_refs), C.byref(h))
        if rc != 0:
            raise RuntimeError(pv._err())
        self.h = h

    def gfa_stats(self):
        """Size of the GFA text and the wall time of (file -> HBM, device parse) of the open."""
        n, ms = C.c_uint64(), (C.c_double * 2)()
        self.lib.pvgpu_session_gfa_stats(self.h, C.byref(n), ms)
        return {"text_bytes": int(n.value), "read_ms": ms[0], "parse_ms": ms[1]}

    def graph_arrays(self, labels=True):
        """The loaded graph as flat numpy arrays (pvgpu_session_copy_graph)."""
        V, E, H, S = self.sizes()
        out = {"ids": np.empty(V, np.uint32), "seg_len": np.empty(V, np.uint32), "ep1": np.empty(E, np.uint32), "ep2": np.empty(E, np.uint32),
               "step_off": np.empty(H + 1, np.uint32), "steps": np.empty(S, np.uint32), "hap_len": np.empty(H, np.uint32)}
        lab_off = lab = None
        if labels:
            nb = int(self.lib.pvgpu_session_label_bytes(self.h))
            lab_off, lab = np.empty(V + 1, np.uint64), np.empty(nb, np.uint8)
        ptr = lambda a: a.ctypes.data if a is not None and a.size else None
        rc = self.lib.pvgpu_session_copy_graph(self.h, ptr(out["ids"]), ptr(out["seg_len"]), ptr(out["ep1"]), ptr(out["ep2"]), out["step_off"].ctypes.data,
                                               ptr(out["steps"]), ptr(out["hap_len"]), lab_off.ctypes.data if labels else None, ptr(lab))
        if rc != 0:
            raise RuntimeError(self.pv._err())
        if labels:
            out["label_off"], out["labels"] = lab_off, lab.tobytes()
        out["path_names"] = [self.lib.pvgpu_session_path_name(self.h, k).decode() for k in range(H)]
        return out

    @classmethod
    def from_graph(cls, pv, g, with_labels=True):
        """Session over a synth/loader graph dict (arrays already in memory; the povu_ffi builder path)."""
        ids = np.ascontiguousarray(g["ids"], dtype=np.uint32)
        V = len(ids)
        ep1, ep2 = edge_ports(g)
        seqs = g.get("seqs")
        lens = np.array([len(x) for x in seqs], dtype=np.uint32) if seqs is not None else np.ones(V, np.uint32)
        lab_off = lab = None
        if with_labels and seqs is not None:
            lab_off = np.concatenate([[0], np.cumsum(lens, dtype=np.uint64)]).astype(np.uint64)
            lab = "".join(seqs).encode()
        paths = g.get("paths", [])
        H = len(paths)
        names = (C.c_char_p * max(H, 1))(*[p[0].encode() for p in paths])
        off = np.zeros(H + 1, np.uint32)
        for i, p in enumerate(paths):
            off[i + 1] = off[i] + len(p[1])
        steps = np.empty(int(off[-1]), np.uint32)
        for i, p in enumerate(paths):
            steps[off[i]:off[i + 1]] = (np.asarray(p[1], np.uint32) << 1) | np.asarray(p[2], np.uint32)
        h = C.c_void_p()
        rc = pv.lib.pvgpu_session_from_arrays(pv.h, V, ids.ctypes.data, lens.ctypes.data, lab_off.ctypes.data if lab_off is not None else None,
                                              lab, len(ep1), ep1.ctypes.data, ep2.ctypes.data, H, names, off.ctypes.data, steps.ctypes.data, C.byref(h))
        if rc != 0:
            raise RuntimeError(pv._err())
        s = cls(pv, _handle=h)
        s.n_steps = int(off[-1])
        return s

    @classmethod
    def from_flat(cls, pv, f, with_labels=True):
        """Session over the flat arrays of synth.tiled_flat (no per-segment Python objects: chromosome-scale components)."""
        ids = np.ascontiguousarray(f["ids"], dtype=np.uint32)
        V = len(ids)
        ep1 = np.ascontiguousarray(f["v1"].astype(np.uint32) * 2 + f["s1"].astype(np.uint32), dtype=np.uint32)
        ep2 = np.ascontiguousarray(f["v2"].astype(np.uint32) * 2 + f["s2"].astype(np.uint32), dtype=np.uint32)
        lens = np.ascontiguousarray(f["seg_len"], dtype=np.uint32)
        lab_off = np.ascontiguousarray(f["lab_off"], dtype=np.uint64) if with_labels else None
        lab = np.ascontiguousarray(f["lab"], dtype=np.uint8) if with_labels else None
        H = len(f["names"])
        if int(f["step_off"][-1]) >= 0xFFFFFFF0:
            raise RuntimeError("more than 2^32 path steps in one session")
        names = (C.c_char_p * max(H, 1))(*[n.encode() for n in f["names"]])
        off = np.ascontiguousarray(f["step_off"], dtype=np.uint32)
        steps = np.ascontiguousarray(f["steps"], dtype=np.uint32)
        h = C.c_void_p()
        rc = pv.lib.pvgpu_session_from_arrays(pv.h, V, ids.ctypes.data, lens.ctypes.data, lab_off.ctypes.data if lab_off is not None else None,
                                              lab.ctypes.data if lab is not None else None, len(ep1), ep1.ctypes.data, ep2.ctypes.data, H, names,
                                              off.ctypes.data, steps.ctypes.data, C.byref(h))
        if rc != 0:
            raise RuntimeError(pv._err())
        s = cls(pv, _handle=h)
        s.n_steps = int(off[-1])
        return s

    def hot_path(self, prefixes, keep_resident=True):
        """decompose -> device index -> call, once.  Returns ms of (upload, decompose, index, call)."""
        arr = (C.c_char_p * len(prefixes))(*[p.encode() for p in prefixes])
        ms = (C.c_double * 4)()
        self._check(self.lib.pvgpu_session_hot_path(self.h, arr, len(prefixes), int(keep_resident), ms))
        return [float(x) for x in ms]

    def close(self):
        if self.h:
            self.lib.pvgpu_session_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.pv._err())

    def _text(self, p, n):
        if not p:
            raise RuntimeError(self.pv._err())
        try:
            return C.string_at(p, n.value).decode()
        finally:
            self.lib.pvgpu_free(p)

    def sizes(self):
        a = [C.c_uint64() for _ in range(4)]
        self.lib.pvgpu_session_sizes(self.h, *[C.byref(x) for x in a])
        return tuple(int(x.value) for x in a)

    def decompose(self):
        self._check(self.lib.pvgpu_session_decompose(self.h))
        return int(self.lib.pvgpu_session_flubble_count(self.h))

    def pvst_texts(self):
        out = {}
        for k in range(self.lib.pvgpu_session_tree_count(self.h)):
            n = C.c_size_t()
            out[int(self.lib.pvgpu_session_tree_component(self.h, k))] = self._text(self.lib.pvgpu_session_pvst_text(self.h, k, C.byref(n)), n)
        return out

    def write_pvst(self, d):
        self._check(self.lib.pvgpu_session_write_pvst(self.h, os.fsencode(d)))

    def read_pvst(self, d):
        self._check(self.lib.pvgpu_session_read_pvst(self.h, os.fsencode(d)))

    def call(self, prefixes):
        arr = (C.c_char_p * len(prefixes))(*[p.encode() for p in prefixes])
        self._check(self.lib.pvgpu_session_call(self.h, arr, len(prefixes)))
        return int(self.lib.pvgpu_session_record_count(self.h))

    def set_restrict(self, region):
        """`povu call -g ref:start-end` (rov.cpp:51-183); None clears the filter."""
        self._check(self.lib.pvgpu_session_set_restrict(self.h, region.encode() if region is not None else None))

    def call_stats(self):
        a = (C.c_uint64 * 12)()
        self.lib.pvgpu_session_call_stats(self.h, a)
        k = ["selected", "regions", "tangled", "pin_pairs", "sort_failed", "hard_errors", "chunks", "too_small", "max_region", "big_bfs", "subr_clamped", "slice_tree_replays"]
        return dict(zip(k, [int(x) for x in a]))

    def vcf_text(self, chunk_size=100, with_date=False, prefix=None):
        n = C.c_size_t()
        return self._text(self.lib.pvgpu_session_vcf_text(self.h, chunk_size, int(with_date), prefix.encode() if prefix else None, C.byref(n)), n)

    def vcf_write(self, path, chunk_size=100, with_date=False, prefix=None):
        """Header + device-written record text straight to a file; returns the number of bytes."""
        n = C.c_uint64()
        fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
        try:
            self._check(self.lib.pvgpu_session_vcf_write(self.h, chunk_size, int(with_date), prefix.encode() if prefix else None, fd, C.byref(n)))
        finally:
            os.close(fd)
        return int(n.value)

    def shard(self, rank, world):
        """Keep only this rank's share of the components (balanced by edges + path steps); returns the total
        number of components of the whole graph."""
        n = C.c_uint32()
        self._check(self.lib.pvgpu_session_shard(self.h, rank, world, C.byref(n)))
        return int(n.value)

    def records_packed_bytes(self):
        return int(self.lib.pvgpu_session_records_packed_bytes(self.h))

    def records_pack(self, dst_device_ptr, capacity):
        """The binary records of the last call as one stream in caller-provided device memory; returns the byte count."""
        n = C.c_uint64()
        self._check(self.lib.pvgpu_session_records_pack(self.h, C.c_void_p(dst_device_ptr), int(capacity), C.byref(n)))
        return int(n.value)

    def reshard(self, rank, world, from_host=False):
        """Repeatable shard(): cuts this rank's share out of the whole graph on the device; from_host=True re-uploads the whole
        graph from the page-locked host arrays first.  Returns (components of the whole graph, wall ms)."""
        n, ms = C.c_uint32(), C.c_double()
        self._check(self.lib.pvgpu_session_reshard(self.h, rank, world, int(from_host), C.byref(n), C.byref(ms)))
        return int(n.value), float(ms.value)

    def tree_regions(self):
        """{component id: number of regions that reached the overlay} for this session's PVST trees."""
        k = self.lib.pvgpu_session_tree_count(self.h)
        comp = np.zeros(k, np.uint32); cnt = np.zeros(k, np.uint64)
        if k:
            self.lib.pvgpu_session_tree_regions(self.h, comp.ctypes.data, cnt.ctypes.data)
        return dict(zip(comp.tolist(), cnt.tolist()))

    def record_lines(self):
        """(lines, keys) of all records: keys[i] = (component id, region ordinal inside the component, ref path id)."""
        ln, n = C.c_size_t(), C.c_size_t()
        off, keys = C.c_void_p(), C.c_void_p()
        p = self.lib.pvgpu_session_record_lines(self.h, C.byref(ln), C.byref(off), C.byref(keys), C.byref(n))
        if not p:
            raise RuntimeError(self.pv._err())
        try:
            text = C.string_at(p, ln.value)
            o = np.ctypeslib.as_array(C.cast(off, C.POINTER(C.c_uint64)), (n.value + 1,)).copy()
            k = np.ctypeslib.as_array(C.cast(keys, C.POINTER(C.c_uint32)), (max(n.value, 1) * 3,)).copy()[: n.value * 3].reshape(-1, 3)
        finally:
            self.lib.pvgpu_free(p); self.lib.pvgpu_free(off); self.lib.pvgpu_free(keys)
        return [text[o[i]:o[i + 1]].decode() for i in range(n.value)], k
