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
