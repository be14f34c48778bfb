"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE: the CPU restatement of the reference).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        vp = C.c_void_p
        _lib.oracle_decompose_arrays.restype = vp
        _lib.oracle_decompose_arrays.argtypes = [C.c_uint32, vp, C.c_uint64, vp, vp, vp, vp]
        _lib.oracle_decompose_gfa.restype = vp
        _lib.oracle_decompose_gfa.argtypes = [C.c_char_p]
        _lib.oracle_result_flubble_count.restype = C.c_uint64
        _lib.oracle_result_flubble_count.argtypes = [vp]
        _lib.oracle_result_component_count.restype = C.c_uint64
        _lib.oracle_result_component_count.argtypes = [vp]
        _lib.oracle_result_copy.argtypes = [vp] * 7
        _lib.oracle_result_copy_components.argtypes = [vp] * 4
        _lib.oracle_result_free.argtypes = [vp]
        _lib.oracle_class_stack_gfa.restype = C.c_int64
        _lib.oracle_class_stack_gfa.argtypes = [C.c_char_p, C.c_uint32, C.c_uint64, vp, vp, vp, vp]
        _lib.oracle_decompose_arrays_notips.restype = vp
        _lib.oracle_decompose_arrays_notips.argtypes = [C.c_uint32, vp, C.c_uint64, vp, vp, vp, vp]
    return _lib


def _rows(r):
    L = lib()
    n = L.oracle_result_flubble_count(r)
    comp = np.zeros(n, np.uint32); a = np.zeros(n, np.uint32); z = np.zeros(n, np.uint32); p = np.zeros(n, np.uint32)
    af = np.zeros(n, np.uint8); zf = np.zeros(n, np.uint8)
    L.oracle_result_copy(r, comp.ctypes.data, a.ctypes.data, af.ctypes.data, z.ctypes.data, zf.ctypes.data, p.ctypes.data)
    L.oracle_result_free(r)
    return list(zip(comp.tolist(), a.tolist(), af.tolist(), z.tolist(), zf.tolist(), p.tolist()))


def oracle_decompose(g, infer_tips=True):
    """Rows (component_id, a_id, a_fwd, z_id, z_fwd, parent_pvst_idx) in creation order."""
    ids = np.ascontiguousarray(g["ids"], dtype=np.uint32)
    v1 = np.ascontiguousarray(g["v1"], dtype=np.uint32); v2 = np.ascontiguousarray(g["v2"], dtype=np.uint32)
    s1 = np.ascontiguousarray(g["s1"], dtype=np.uint8); s2 = np.ascontiguousarray(g["s2"], dtype=np.uint8)
    fn = lib().oracle_decompose_arrays if infer_tips else lib().oracle_decompose_arrays_notips
    r = fn(len(ids), ids.ctypes.data, len(v1), v1.ctypes.data, s1.ctypes.data, v2.ctypes.data, s2.ctypes.data)
    assert r, "oracle failed"
    return _rows(r)


def flubble_rows(f, ids):
    """Same row format from a povu_b200.api.Flubbles."""
    return list(zip((f.comp + 1).tolist(), ids[f.a_v].tolist(), f.a_fwd.tolist(), ids[f.z_v].tolist(), f.z_fwd.tolist(), f.parent.tolist()))


def oracle_decompose_gfa(path):
    r = lib().oracle_decompose_gfa(path.encode())
    assert r, "oracle failed"
    return _rows(r)


def oracle_class_stack(path, k=0, cap=1 << 22):
    e = np.zeros(cap, np.uint32); c = np.zeros(cap, np.uint32); v = np.zeros(cap, np.uint32); f = np.zeros(cap, np.uint8)
    n = lib().oracle_class_stack_gfa(path.encode(), k, cap, e.ctypes.data, c.ctypes.data, v.ctypes.data, f.ctypes.data)
    assert n >= 0
    return e[:n], c[:n], v[:n], f[:n]


def rows_to_pvst_texts(rows):
    """{component_id: .pvst text} from oracle/GPU rows (to_pvst.cpp:23-109)."""
    by = {}
    for comp, a, af, z, zf, p in rows:
        by.setdefault(comp, []).append((f"{'>' if af else '<'}{a}{'>' if zf else '<'}{z}", p))
    out = {}
    for comp, fl in by.items():
        ch = [[] for _ in range(len(fl) + 1)]
        for j, (_, p) in enumerate(fl):
            ch[p].append(j + 1)
        lines = ["H\t0.0.3\t.\t.\t.\n"]
        for i in range(len(fl) + 1):
            kids = ", ".join(map(str, ch[i])) if ch[i] else "."
            lines.append(f"D\t0\t.\t{kids}\t.\n" if i == 0 else f"F\t{i}\t{fl[i - 1][0]}\t{kids}\tL\n")
        out[comp] = "".join(lines)
    return out
