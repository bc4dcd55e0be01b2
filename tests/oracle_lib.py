"""TEST INFRASTRUCTURE: ctypes binding of oracle/liboracle.so (the CPU restatement) and a runner for
oracle/_ref/ref_overlap (the unmodified reference behind oracle/ref_harness.cpp)."""
import ctypes as C
import json
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_overlap")

_lib = None


def _olib():
    global _lib
    if _lib is None:
        l = C.CDLL(ORACLE_SO)
        vp, u64 = C.c_void_p, C.c_uint64
        l.oracle_create.restype = vp
        l.oracle_destroy.argtypes = [vp]
        l.oracle_load_reads.argtypes = [vp, vp, vp, u64, C.c_uint32]
        for f in ("oracle_n_unique", "oracle_n_good", "oracle_shortest", "oracle_longest", "oracle_total_bases"):
            getattr(l, f).restype = u64
            getattr(l, f).argtypes = [vp]
        l.oracle_read_info.argtypes = [vp, vp, vp, vp, vp]
        l.oracle_get_read.argtypes = [vp, u64, C.c_int, vp]
        l.oracle_sorted_reads.argtypes = [vp, vp, vp]
        l.oracle_build_index.argtypes = [vp]
        l.oracle_hash_string_length.argtypes = [vp]
        l.oracle_hash_string_length.restype = C.c_uint32
        l.oracle_lookup.argtypes = [vp, C.c_char_p, C.c_uint32, vp, C.c_uint32]
        l.oracle_lookup.restype = C.c_uint32
        l.oracle_mark_contained.argtypes = [vp]
        l.oracle_build_graph.argtypes = [vp, C.c_int, C.c_int]
        l.oracle_n_edges.argtypes = [vp, C.c_int]
        l.oracle_n_edges.restype = u64
        l.oracle_get_edges.argtypes = [vp, C.c_int, vp]
        l.oracle_counters.argtypes = [vp, vp]
        _lib = l
    return _lib


class Oracle:
    """CPU restatement of Dataset -> HashTable -> OverlapGraph (see oracle/omega_oracle.cpp)."""
    BFS, THREE_PHASE = 0, 1

    def __init__(self, bases, offsets, min_overlap):
        self.l = _olib()
        self.h = self.l.oracle_create()
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.l.oracle_load_reads(self.h, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1, min_overlap)
        self.n = self.l.oracle_n_unique(self.h)
        self.n_good = self.l.oracle_n_good(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.l.oracle_destroy(self.h)
            self.h = None

    def read_info(self):
        sup = np.zeros(self.n, np.uint64); ln = np.zeros(self.n, np.uint32)
        fr = np.zeros(self.n, np.uint32); fnv = np.zeros(self.n, np.uint64)
        self.l.oracle_read_info(self.h, sup.ctypes.data, ln.ctypes.data, fr.ctypes.data, fnv.ctypes.data)
        return dict(sup=sup, len=ln, freq=fr, fnv=fnv)

    def get_read(self, rid, reverse=False):
        buf = C.create_string_buffer(65536)
        n = self.l.oracle_get_read(self.h, rid, 1 if reverse else 0, buf)
        return buf.raw[:n].decode()

    def sorted_reads(self):
        tot = self.l.oracle_total_bases(self.h)
        bases = np.zeros(tot, np.uint8); offs = np.zeros(self.n + 1, np.uint64)
        self.l.oracle_sorted_reads(self.h, bases.ctypes.data, offs.ctypes.data)
        return bases, offs

    def build_index(self):
        self.l.oracle_build_index(self.h)
        return self.l.oracle_hash_string_length(self.h)

    def lookup(self, key):
        k = key.encode() if isinstance(key, str) else bytes(key)
        cap = 64
        while True:
            out = np.zeros(cap, np.uint64)
            n = self.l.oracle_lookup(self.h, k, len(k), out.ctypes.data, cap)
            if n <= cap:
                return out[:n]
            cap = n

    def mark_contained(self):
        self.l.oracle_mark_contained(self.h)

    def build_graph(self, mode=0, threads=1):
        self.l.oracle_build_graph(self.h, mode, threads)

    def edges(self, pre=False):
        n = self.l.oracle_n_edges(self.h, 1 if pre else 0)
        out = np.zeros((n, 4), np.uint32)
        self.l.oracle_get_edges(self.h, 1 if pre else 0, out.ctypes.data)
        return out

    def counters(self):
        out = np.zeros(10, np.uint64)
        self.l.oracle_counters(self.h, out.ctypes.data)
        keys = ["number_of_nodes", "number_of_edges", "P_c", "P_e", "C_c", "C_e", "T", "active_pivots", "max_degree", "E_pre"]
        return dict(zip(keys, (int(x) for x in out)))

    def run_all(self, mode=0, threads=1):
        self.build_index()
        self.mark_contained()
        self.build_graph(mode, threads)
        return self


LEAN_SO = os.path.join(ROOT, "oracle", "liblean.so")
_lean = None


def _llib():
    global _lean
    if _lean is None:
        l = C.CDLL(LEAN_SO)
        vp, u64 = C.c_void_p, C.c_uint64
        l.lean_create.restype = vp
        l.lean_create.argtypes = [C.c_int]
        l.lean_destroy.argtypes = [vp]
        l.lean_load_reads.argtypes = [vp, vp, vp, u64, C.c_uint32]
        for f in ("lean_n_unique", "lean_n_good", "lean_n_edges"):
            getattr(l, f).restype = u64
            getattr(l, f).argtypes = [vp]
        l.lean_run.argtypes = [vp, C.c_int]
        l.lean_read_info.argtypes = [vp, vp, vp, vp, vp]
        l.lean_counters.argtypes = [vp, vp]
        l.lean_get_edges.argtypes = [vp, vp]
        l.lean_degrees.argtypes = [vp, vp]
        _lean = l
    return _lean


class LeanOracle:
    """Memory-lean three-phase restatement (oracle/lean_oracle.cpp) for BASELINE.json sizes: counters and an
    order-independent checksum of the final edge set; the tuples themselves only when keep_edges is set."""

    def __init__(self, bases, offsets, min_overlap, threads=None):
        self.l = _llib()
        self.h = self.l.lean_create(threads or os.cpu_count() or 1)
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.l.lean_load_reads(self.h, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1, min_overlap)
        self.n = self.l.lean_n_unique(self.h)
        self.n_good = self.l.lean_n_good(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.l.lean_destroy(self.h)
            self.h = None

    def run(self, keep_edges=False):
        if self.l.lean_run(self.h, 1 if keep_edges else 0) != 0:
            raise RuntimeError("lean oracle: hashStringLength must be 1..64")
        return self

    def counters(self):
        out = np.zeros(14, np.uint64)
        self.l.lean_counters(self.h, out.ctypes.data)
        keys = ["n_unique", "E_pre", "E_final", "nodes", "contained", "max_degree", "P_e", "T", "active_pivots", "P_c", "C_c",
                "ck_xor", "ck_sum", "asymmetric"]
        return dict(zip(keys, (int(x) for x in out)))

    def checksum(self):
        c = self.counters()
        return [c["ck_xor"], c["ck_sum"]]

    def edges(self):
        n = self.l.lean_n_edges(self.h)
        out = np.zeros((n, 4), np.uint32)
        self.l.lean_get_edges(self.h, out.ctypes.data)
        return out

    def read_info(self):
        sup = np.zeros(self.n, np.uint64); ln = np.zeros(self.n, np.uint32)
        fr = np.zeros(self.n, np.uint32); fnv = np.zeros(self.n, np.uint64)
        self.l.lean_read_info(self.h, sup.ctypes.data, ln.ctypes.data, fr.ctypes.data, fnv.ctypes.data)
        return dict(sup=sup, len=ln, freq=fr, fnv=fnv)

    def degrees(self):
        out = np.zeros(self.n, np.uint32)
        self.l.lean_degrees(self.h, out.ctypes.data)
        return out


def edge_checksum(e):
    """xor and sum of a 64-bit mix of every (src, dst, offset, orient) tuple: independent of order
    (the same mix as lean_oracle.cpp mix_tuple and tests/golden/make_full_size.py)."""
    e = np.asarray(e).astype(np.uint64).reshape(-1, 4)
    x = (e[:, 0] * np.uint64(0x9E3779B97F4A7C15) ^ e[:, 1] * np.uint64(0xC2B2AE3D27D4EB4F) ^ e[:, 2] * np.uint64(0x165667B19E3779F9)
         ^ e[:, 3] * np.uint64(0x27D4EB2F165667C5))
    x ^= x >> np.uint64(29); x *= np.uint64(0xBF58476D1CE4E5B9); x ^= x >> np.uint64(32)
    return [int(np.bitwise_xor.reduce(x)) if len(x) else 0, int(x.sum(dtype=np.uint64))]


def have_reference():
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


DUMP_READ = np.dtype([("sup", "<u8"), ("len", "<u4"), ("freq", "<u4"), ("fnv", "<u8")])


def read_dump(path):
    with open(path, "rb") as f:
        hdr = np.frombuffer(f.read(48), dtype="<u8")
        assert hdr[0] == 0x31504d554442474f, "bad dump magic"
        n, ne = int(hdr[1]), int(hdr[2])
        reads = np.frombuffer(f.read(n * DUMP_READ.itemsize), dtype=DUMP_READ)
        edges = np.frombuffer(f.read(ne * 16), dtype="<u4").reshape(ne, 4)
    return dict(n=n, number_of_nodes=int(hdr[3]), number_of_edges=int(hdr[4]), h=int(hdr[5]), reads=reads,
                edges=sort_tuples(edges))


def read_mates(path):
    """Mate-pair lists dumped by oracle/ref_harness.cpp --mates or metagenomics_b200/host/ogb_overlap --mates:
    (start (n+1,) int64, triples (m,3) uint32 [matePairID, matePairOrientation, datasetNumber]) in list order, reads 1..n."""
    raw = np.fromfile(path, dtype=np.uint8)
    hdr = np.frombuffer(raw[:16].tobytes(), dtype="<u8")
    assert hdr[0] == 0x31534554414d474f, "bad mates magic"
    n = int(hdr[1])
    w = np.frombuffer(raw[16:].tobytes(), dtype="<u4")
    start = np.zeros(n + 1, np.int64)
    out, p = [], 0
    for i in range(n):
        c = int(w[p]); p += 1
        out.append(w[p:p + 3 * c].reshape(c, 3)); p += 3 * c
        start[i + 1] = start[i] + c
    assert p == len(w)
    return start, (np.concatenate(out) if out else np.zeros((0, 3), np.uint32))


def read_dump2(path):
    """Second dump of oracle/ref_harness.cpp (--dump2): the reference graph after its contractCompositePaths /
    removeDeadEndNodes fix-point (OverlapGraph.cpp:211-215). Returns edges (n,5) [src,dst,orient,nlist,offset] (uint64),
    and the concatenated per-edge lists (m,3) [read, overlapOffset, orientation] with list_start (n+1)."""
    raw = np.fromfile(path, dtype=np.uint8)
    hdr = np.frombuffer(raw[:40].tobytes(), dtype="<u8")
    assert hdr[0] == 0x32504d554442474f, "bad dump2 magic"
    n, ne = int(hdr[1]), int(hdr[2])
    edges = np.zeros((ne, 5), np.uint64); starts = np.zeros(ne + 1, np.int64); lists = []
    p = 40
    for i in range(ne):
        a = np.frombuffer(raw[p:p + 16].tobytes(), dtype="<u4"); off = int(np.frombuffer(raw[p + 16:p + 24].tobytes(), dtype="<u8")[0]); p += 24
        nl = int(a[3])
        edges[i] = (a[0], a[1], a[2], nl, off)
        lists.append(np.frombuffer(raw[p:p + 12 * nl].tobytes(), dtype="<u4").reshape(nl, 3)); p += 12 * nl
        starts[i + 1] = starts[i] + nl
    assert p == len(raw)
    return dict(n=n, number_of_nodes=int(hdr[3]), number_of_edges=int(hdr[4]), edges=edges, list_start=starts,
                lists=np.concatenate(lists) if lists else np.zeros((0, 3), np.uint32))


def sort_tuples(e):
    """canonical order (src, offset, dst, orient) of an (n,4) [src,dst,offset,orient] array."""
    e = np.asarray(e, dtype=np.uint32).reshape(-1, 4)
    if len(e) == 0:
        return e
    idx = np.lexsort((e[:, 3], e[:, 1], e[:, 2], e[:, 0]))
    return np.ascontiguousarray(e[idx])


def run_reference(fasta_paths, min_overlap, paired=False, want_table=False, binary=REF_BIN, timeout=3600, contracted=False, unitig=None, mates=False):
    """Runs the unmodified reference on FASTA files; returns (dump dict, timing json, table lists|None). contracted=True:
    the dump dict gets a "contracted" entry (read_dump2: the graph after OverlapGraph.cpp:211-215). unitig=path (with
    contracted): the reference's own sortEdges + saveGraphToFile (main.cpp:49-50) writes its .unitig file there."""
    with tempfile.TemporaryDirectory() as td:
        dump, js, tab = os.path.join(td, "d.bin"), os.path.join(td, "t.json"), os.path.join(td, "tab.bin")
        cmd = [binary, "-l", str(min_overlap), "--dump", dump, "--json", js]
        if contracted:
            cmd += ["--dump2", os.path.join(td, "d2.bin")]
            if unitig:
                cmd += ["--unitig", os.path.abspath(unitig)]
        for p in fasta_paths:
            cmd += ["-pe" if paired else "-se", p]
        if want_table:
            cmd += ["--table", tab]
        if mates:
            cmd += ["--mates", os.path.join(td, "m.bin")]
        subprocess.run(cmd, check=True, timeout=timeout, cwd=td)
        d = read_dump(dump)
        if mates:
            d["mates"] = read_mates(os.path.join(td, "m.bin"))
        if contracted:
            d["contracted"] = read_dump2(os.path.join(td, "d2.bin"))
        with open(js) as f:
            t = json.load(f)
        table = None
        if want_table:
            raw = np.fromfile(tab, dtype=np.uint8)
            table, p = [], 0
            for _ in range(d["n"] * 4):
                c = int(np.frombuffer(raw[p:p + 4].tobytes(), "<u4")[0]); p += 4
                table.append(np.frombuffer(raw[p:p + 8 * c].tobytes(), "<u8").copy()); p += 8 * c
        return d, t, table
