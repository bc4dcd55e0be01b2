"""Helpers of the simplification tests (tests/test_contract.py on the CPU, tests/test_gpu_parity.py on the GPU)."""
import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

EDGE = np.dtype([("src", "<u4"), ("dst", "<u4"), ("offset", "<u2"), ("orient", "u1"), ("reserved", "u1")])
CEDGE = np.dtype([("src", "<u4"), ("dst", "<u4"), ("offset", "<u8"), ("list_start", "<u8"), ("count", "<u4"), ("twin", "<u4"), ("orient", "u1"),
                  ("reserved", "u1", (3,)), ("reserved2", "<u4")])
ITEM = np.dtype([("read", "<u4"), ("offset", "<u2"), ("orient", "u1"), ("reserved", "u1")])


def load_oracle_module(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "oracle", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def fixture_records(g):
    ce, ls, cl = g["c_edges"].astype(np.int64), g["c_list_start"], g["c_lists"].astype(np.int64)
    return sorted((int(s), int(d), int(o), int(off), tuple(cl[ls[i]:ls[i + 1], 0].tolist()), tuple(cl[ls[i]:ls[i + 1], 1].tolist()),
                   tuple(cl[ls[i]:ls[i + 1], 2].tolist())) for i, (s, d, o, nl, off) in enumerate(ce.tolist()))


def composite_records(edges, items):
    """(ogb_cedge array, ogb_clist_item array) -> the sorted record tuples of the oracles' edge_records()."""
    out = []
    for e in edges:
        a, b = int(e["list_start"]), int(e["list_start"]) + int(e["count"])
        out.append((int(e["src"]), int(e["dst"]), int(e["orient"]), int(e["offset"]), tuple(items["read"][a:b].tolist()),
                    tuple(items["offset"][a:b].tolist()), tuple(items["orient"][a:b].tolist())))
    return sorted(out)


def check_twins(edges):
    """twin links of the result: an involution onto the reversed end points with the twin orientation"""
    t = edges["twin"].astype(np.int64)
    assert (t[t] == np.arange(len(edges))).all()
    assert (edges["src"][t] == edges["dst"]).all() and (edges["dst"][t] == edges["src"]).all()
    tw = np.array([3, 1, 2, 0])
    assert (edges["orient"][t] == tw[edges["orient"]]).all()
    assert (edges["count"][t] == edges["count"]).all()


def edges_struct(tuples):
    """(n, 4) [src, dst, offset, orient] -> ogb_edge records in the library's order (src, offset, dst, orient)"""
    t = np.asarray(tuples, dtype=np.int64).reshape(-1, 4)
    order = np.lexsort((t[:, 3], t[:, 1], t[:, 2], t[:, 0]))
    t = t[order]
    e = np.zeros(len(t), dtype=EDGE)
    e["src"], e["dst"], e["offset"], e["orient"] = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
    return e



def unitig_records(path):
    """A .unitig file (OverlapGraph::saveGraphToFile, OverlapGraph.cpp:1219-1259) -> sorted record tuples like edge_records()."""
    v = [int(x) for x in open(path).read().split()]
    out, i = [], 0
    while i < len(v):
        s, d, o, off, k = v[i:i + 5]
        i += 5
        body = v[i:i + 3 * k]
        i += 3 * k
        out.append((s, d, o, off, tuple(body[0::3]), tuple(body[1::3]), tuple(body[2::3])))
    return sorted(out)


def build_emul():
    """Compiles tests/contract_emul.cpp (the kernel bodies of csrc/ogb_contract.cuh on the CPU) and returns run(edges, lens, reverse)."""
    out = os.path.join(HERE, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libcontract_emul.so")
    src = os.path.join(HERE, "contract_emul.cpp")
    hdr = os.path.join(ROOT, "metagenomics_b200", "csrc", "ogb_contract.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", so], check=True)
    lib = C.CDLL(so)
    lib.emul_simplify.restype = C.c_int
    lib.emul_simplify.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
                                  C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_void_p]
    lib.emul_free.argtypes = [C.c_void_p]

    def run(edge_tuples, lens, reverse=False, presorted=None):
        e = presorted if presorted is not None else edges_struct(edge_tuples)
        lens = np.ascontiguousarray(lens, dtype=np.uint16)
        pe, pi, ne, ni = C.c_void_p(), C.c_void_p(), C.c_uint64(), C.c_uint64()
        stats = np.zeros(5, dtype=np.uint64)
        rc = lib.emul_simplify(e.ctypes.data, len(e), lens.ctypes.data, len(lens), int(reverse), C.byref(pe), C.byref(ne), C.byref(pi), C.byref(ni), stats.ctypes.data)
        assert rc == 0, rc
        edges = np.frombuffer(C.string_at(pe.value, ne.value * CEDGE.itemsize), dtype=CEDGE).copy()
        items = np.frombuffer(C.string_at(pi.value, ni.value * ITEM.itemsize), dtype=ITEM).copy()
        lib.emul_free(pe); lib.emul_free(pi)
        return edges, items, stats
    return run


# ---- the sequential C++ restatement (oracle/contract_seq.cpp) and the checksum of a simplified graph ----

class _SeqResult(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_edges", "n_items", "merges", "dead_ends", "iterations", "ck_xor", "ck_sum")]


def seq_simplify(edge_tuples, lens, arrays=True):
    """oracle/libcontractseq.so: the reference's fix-point in its own sequential order. Returns (stats dict, edges, items): edges
    (n, 6) uint64 [src, dst, orient, offset, count, list_start], items (m, 3) uint32 [read, offset, orientation] (None, None
    without arrays); stats carries the order-independent checksum [ck_xor, ck_sum]."""
    lib = C.CDLL(os.path.join(ROOT, "oracle", "libcontractseq.so"))
    lib.cseq_simplify.restype = C.c_int
    lib.cseq_simplify.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.POINTER(_SeqResult), C.c_void_p, C.c_void_p]
    lib.cseq_free.argtypes = [C.c_void_p]
    e = np.ascontiguousarray(np.asarray(edge_tuples).reshape(-1, 4), dtype=np.uint32)
    lens = np.ascontiguousarray(lens, dtype=np.uint16)
    res, pe, pi = _SeqResult(), C.c_void_p(), C.c_void_p()
    rc = lib.cseq_simplify(e.ctypes.data, len(e), lens.ctypes.data, len(lens), C.byref(res), C.byref(pe) if arrays else None, C.byref(pi) if arrays else None)
    assert rc == 0, rc
    st = {n: int(getattr(res, n)) for n, _ in _SeqResult._fields_}
    st["checksum"] = [st.pop("ck_xor"), st.pop("ck_sum")]
    if not arrays:
        return st, None, None
    edges = np.frombuffer(C.string_at(pe.value, st["n_edges"] * 48), dtype=np.uint64).reshape(-1, 6).copy()
    items = np.frombuffer(C.string_at(pi.value, st["n_items"] * 12), dtype=np.uint32).reshape(-1, 3).copy()
    lib.cseq_free(pe); lib.cseq_free(pi)
    return st, edges, items


def seq_records(edges, items):
    out = []
    for s, d, o, off, k, a in edges.tolist():
        out.append((s, d, o, off, tuple(items[a:a + k, 0].tolist()), tuple(items[a:a + k, 1].tolist()), tuple(items[a:a + k, 2].tolist())))
    return sorted(out)


def _mix(x):
    x = x ^ (x >> np.uint64(29))
    x = x * np.uint64(0xBF58476D1CE4E5B9)
    return x ^ (x >> np.uint64(32))


def simplified_checksum(edges, items):
    """The figure of oracle/contract_seq.cpp (checksum_edge / checksum_item) for a simplified graph given as ogb_cedge /
    ogb_clist_item arrays (metagenomics_b200.api CEDGE_DTYPE / CITEM_DTYPE): independent of the order of the edges, dependent
    on the order inside every list. Lists must be laid out in edge order (list_start ascending), as the library returns them."""
    K1, K2, K3, K4, K5 = (np.uint64(k) for k in (0x9E3779B97F4A7C15, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9, 0x27D4EB2F165667C5, 0x94D049BB133111EB))
    with np.errstate(over="ignore"):
        cnt = edges["count"].astype(np.uint64)
        start = edges["list_start"].astype(np.int64)
        assert len(items) == int(cnt.sum()) and (len(edges) == 0 or (np.array_equal(start, np.concatenate([[0], np.cumsum(cnt.astype(np.int64))[:-1]]))))
        pos = np.arange(len(items), dtype=np.int64) - np.repeat(start, cnt.astype(np.int64)) + 1
        h = _mix(items["read"].astype(np.uint64) * K1 ^ items["offset"].astype(np.uint64) * K2 ^ items["orient"].astype(np.uint64) * K3 ^ pos.astype(np.uint64) * K4)
        csum = np.concatenate([[np.uint64(0)], np.cumsum(h, dtype=np.uint64)])
        list_sum = csum[start + cnt.astype(np.int64)] - csum[start]
        eh = _mix(edges["src"].astype(np.uint64) * K1 ^ edges["dst"].astype(np.uint64) * K2 ^ edges["offset"].astype(np.uint64) * K3
                  ^ edges["orient"].astype(np.uint64) * K4 ^ cnt * K5)
        tot = _mix(eh + list_sum)
        return [int(np.bitwise_xor.reduce(tot)) if len(tot) else 0, int(tot.sum(dtype=np.uint64))]
