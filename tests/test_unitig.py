"""CPU: the reference's .unitig text format (OverlapGraph::saveGraphToFile / readGraphFromFile, OverlapGraph.cpp:1219-1367;
resume path main.cpp:36-42) in the C++ drop-in classes. Fixtures: the files the UNMODIFIED reference wrote for the golden
data sets (tests/golden/*.unitig, produced by tests/golden/make_golden.py through oracle/ref_harness.cpp --unitig, i.e. the
reference's own sortEdges + saveGraphToFile after its contraction fix-point): composite edges with read / offset /
orientation lists, palindromic self-edges. The drop-in's reader must rebuild exactly that graph -- every edge with its
reverse edge -- and its writer must reproduce the file byte for byte. No GPU involved (the resume path never builds)."""
import glob
import os
import subprocess

import numpy as np
import pytest

from oracle_lib import read_dump, sort_tuples

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
EXE = os.path.join(HERE, "..", "metagenomics_b200", "host", "ogb_overlap")
NAMES = sorted(os.path.basename(p)[:-7] for p in glob.glob(os.path.join(GOLDEN, "*.unitig")))


def test_unitig_fixtures_present():
    assert len(NAMES) >= 6


@pytest.mark.parametrize("name", NAMES)
def test_reader_and_writer_reproduce_the_reference_file(name, tmp_path):
    from metagenomics_b200 import synth
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    fa = str(tmp_path / "in.fa")
    synth.write_fasta(fa, z["bases"], z["offsets"])
    prefix = str(tmp_path / "g")
    want = open(os.path.join(GOLDEN, name + ".unitig"), "rb").read()
    open(prefix + ".unitig", "wb").write(want)
    kind = "-pe" if name == "config2_small" else "-se"
    r = subprocess.run([EXE, "-l", str(int(z["min_overlap"])), kind, "1", fa, "-f", prefix, "-s", "--resave", prefix + ".again", "--dump", prefix + ".bin"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0, r.stdout
    assert open(prefix + ".again", "rb").read() == want                       # writer == the reference's writer, byte for byte
    d = read_dump(prefix + ".bin")
    ce = z["c_edges"].astype(np.int64)                                         # (src, dst, orient, nlist, offset) after the reference's contraction
    assert d["n"] == len(z["sup"])
    assert d["number_of_edges"] == int(z["c_number_of_edges"]) and d["number_of_nodes"] == int(z["c_number_of_nodes"])
    assert np.array_equal(d["edges"], sort_tuples(np.stack([ce[:, 0], ce[:, 1], ce[:, 4], ce[:, 2]], axis=1)))   # reader rebuilt both directions
