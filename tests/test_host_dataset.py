"""CPU: the product's host Dataset stage (libogb ogb_dataset_*) against the oracle and the goldens:
filter, case folding, canonical strand, lexicographic order with prefix rule, dedupe/frequency, IDs."""
import glob
import os

import numpy as np
import pytest

import datasets
from oracle_lib import Oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fnv1a(s):
    h = 1469598103934665603
    for c in s.encode():
        h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))), ids=lambda p: os.path.basename(p)[:-4])
def test_dataset_matches_reference_ids(path):
    from metagenomics_b200 import Dataset
    z = np.load(path)
    ds = Dataset(bases=z["bases"], offsets=z["offsets"], minOverlap=int(z["min_overlap"]))
    assert ds.getNumberOfReads() == int(z["n_good"])
    assert ds.getNumberOfUniqueReads() == len(z["fnv"])
    assert np.array_equal(ds.lengths().astype(np.uint32), z["len"])
    assert np.array_equal(ds.frequencies(), z["freq"])
    for i in range(1, len(z["fnv"]) + 1, 7):
        assert fnv1a(ds.getReadFromID(i)) == int(z["fnv"][i - 1])
    assert ds.shortestReadLength == int(z["len"].min()) and ds.longestReadLength == int(z["len"].max())


def test_dataset_strings_and_lookup():
    from metagenomics_b200 import Dataset
    for cfg in (datasets.filtered(), datasets.tandem(mixed=True)):
        ds = Dataset(bases=cfg["bases"], offsets=cfg["offsets"], minOverlap=cfg["min_overlap"])
        orc = Oracle(cfg["bases"], cfg["offsets"], cfg["min_overlap"])
        assert ds.getNumberOfUniqueReads() == orc.n and ds.getNumberOfReads() == orc.n_good
        for i in range(1, orc.n + 1):
            f = orc.get_read(i)
            assert ds.getReadFromID(i) == f and ds.getReadFromID(i, reverse=True) == orc.get_read(i, True)
            if i % 5 == 0:                                        # getReadFromString finds both strands
                assert ds.getReadFromString(f) == i and ds.getReadFromString(datasets.rc(f).lower()) == i
        assert ds.getReadFromString("ACGT" * 20) == 0
        words, offs, lens = ds.packed()
        assert len(offs) == orc.n + 1 and offs[-1] == len(words) == sum((int(l) + 31) // 32 for l in lens)


def test_prefix_orders_first_and_threshold():
    from metagenomics_b200 import Dataset
    a = "ACGTACGTTAGCCGATAGCTAGCTAGGATCGA"
    reads = [a + "AAAC", a, a + "A", a + "AA", "A" * 39 + "C" * 10 + "G", "A" * 40 + "C" * 10]
    ds = Dataset(reads=reads, minOverlap=20)
    orc = Oracle(*__import__("metagenomics_b200.api", fromlist=["x"])._reads_to_buffers(reads), 20)
    got = [ds.getReadFromID(i) for i in range(1, ds.getNumberOfUniqueReads() + 1)]
    want = [orc.get_read(i) for i in range(1, orc.n + 1)]
    assert got == want and got == sorted(got)
    assert ("A" * 40 + "C" * 10) not in got                      # 40 of 50 hits the (UINT64)(len*.8) threshold


def test_fasta_and_fastq_files(tmp_path):
    from metagenomics_b200 import Dataset
    cfg = datasets.repeats()
    b, o = cfg["bases"], cfg["offsets"]
    reads = [bytes(b[int(o[i]):int(o[i + 1])]).decode() for i in range(len(o) - 1)]
    fa, fq = tmp_path / "r.fasta", tmp_path / "r.fastq"
    with open(fa, "w") as f:
        for i, r in enumerate(reads):
            f.write(f">read{i}\n{r[:40]}\n{r[40:]}\n")           # multi-line records are joined (Dataset.cpp:145)
    with open(fq, "w") as f:
        for i, r in enumerate(reads):
            f.write(f"@read{i}\n{r}\n+\n{'I' * len(r)}\n")
    want = Dataset(reads=reads, minOverlap=cfg["min_overlap"])
    for p in (fa, fq):
        ds = Dataset(singleEndFileNames=[str(p)], minOverlap=cfg["min_overlap"])
        assert ds.getNumberOfUniqueReads() == want.getNumberOfUniqueReads()
        assert np.array_equal(ds.frequencies(), want.frequencies())
        assert ds.getReadFromID(17) == want.getReadFromID(17)
    from metagenomics_b200 import OgbError
    with pytest.raises(OgbError):
        Dataset(singleEndFileNames=[str(tmp_path / "missing.fa")], minOverlap=30)
