"""Seeded small data sets shared by the parity tests (all synthetic, no network)."""
import numpy as np

from metagenomics_b200 import synth

COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def rc(s):
    return "".join(COMP[c] for c in reversed(s))


def from_strings(reads, min_overlap, name):
    bs = [r.encode() for r in reads]
    offs = np.zeros(len(bs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    return dict(name=name, bases=np.frombuffer(b"".join(bs), dtype=np.uint8).copy(), offsets=offs, min_overlap=min_overlap, paired=False)


def _rand_seq(rng, n):
    return "".join("ACGT"[i] for i in rng.integers(0, 4, n))


def _sample(rng, genome, n_reads, lmin, lmax):
    out = []
    for _ in range(n_reads):
        L = int(rng.integers(lmin, lmax + 1))
        st = int(rng.integers(0, len(genome) - L + 1))
        s = genome[st:st + L]
        out.append(s if rng.integers(0, 2) else rc(s))
    return out


def repeats(seed=11, min_overlap=30, mixed=False):
    """Exact repeats + inverted repeats: multi-edges, ambiguous pivots."""
    rng = np.random.default_rng(seed)
    unit = _rand_seq(rng, 180)
    g = _rand_seq(rng, 700) + unit + _rand_seq(rng, 500) + unit + _rand_seq(rng, 400) + rc(unit) + _rand_seq(rng, 600)
    lm = (60, 110) if mixed else (80, 80)
    return from_strings(_sample(rng, g, 2500, *lm), min_overlap, f"repeats(mixed={mixed})")


def tandem(seed=12, min_overlap=24, mixed=False):
    """Short-period tandem arrays: self-overlaps, high multiplicity multi-edges, offset ties."""
    rng = np.random.default_rng(seed)
    g = _rand_seq(rng, 400)
    for period, copies in ((7, 30), (12, 20), (31, 9)):
        g += _rand_seq(rng, period) * copies + _rand_seq(rng, 300)
    lm = (50, 90) if mixed else (64, 64)
    return from_strings(_sample(rng, g, 3000, *lm), min_overlap, f"tandem(mixed={mixed})")


def palindromes(seed=13, min_overlap=20):
    """Reads ending in DNA palindromes >= minOverlap: reverse-complement self-overlaps, which the
    reference holds twice (edge + twin object are the same tuple)."""
    rng = np.random.default_rng(seed)
    reads = []
    for _ in range(40):
        half = _rand_seq(rng, int(rng.integers(12, 20)))
        pal = half + rc(half)
        left = _rand_seq(rng, 70)
        full = left + pal + _rand_seq(rng, 70)
        reads.append(full[:70 + len(pal)])                   # ends with the palindrome
        reads.append(full[70:])                              # starts with it
        reads += _sample(rng, full, 12, 60, 60 + len(pal))
    return from_strings(reads, min_overlap, "palindromes")


def even_h(seed=14):
    """minOverlap odd -> even hashStringLength: a key can equal its own reverse complement."""
    rng = np.random.default_rng(seed)
    g = _rand_seq(rng, 3000)
    reads = _sample(rng, g, 1500, 70, 70)
    half = _rand_seq(rng, 15)
    pal = half + rc(half)                                    # 30-mer palindrome = key for h = 30
    reads += [pal + _rand_seq(rng, 40) for _ in range(3)] + [_rand_seq(rng, 40) + pal for _ in range(3)]
    return from_strings(reads, 31, "even_h")


def filtered(seed=15, min_overlap=30):
    """N-containing, low-complexity, too-short, lower-case and duplicated reads (Dataset.cpp:155-167,398-413)."""
    rng = np.random.default_rng(seed)
    g = _rand_seq(rng, 2500)
    reads = _sample(rng, g, 1200, 60, 60)
    reads += [r.lower() for r in reads[:50]]                 # case folding -> duplicates
    reads += [rc(r) for r in reads[50:120]]                  # duplicates on the other strand
    reads += ["A" * 60, "A" * 48 + _rand_seq(rng, 12), "A" * 47 + "CGT" * 4 + "C"]   # 80 % rule: 48/60 is at the threshold
    reads += [r[:20] + "N" + r[21:] for r in reads[200:230]]
    reads += [_rand_seq(rng, 30), _rand_seq(rng, 31), _rand_seq(rng, 29), ""]        # len > minOverlap is strict
    return from_strings(reads, min_overlap, "filtered")


def one_window(seed=16, min_overlap=40):
    """Reads of length minOverlap+1: exactly one window per read (j = 1)."""
    rng = np.random.default_rng(seed)
    g = _rand_seq(rng, 900)
    return from_strings(_sample(rng, g, 3000, min_overlap + 1, min_overlap + 1), min_overlap, "one_window")


def primer_prefixed(seed=17, min_overlap=30, n_reads=24000):
    """Amplicon-like reads: every read starts with the same 20-base primer, so one of its four index keys has the same
    leading bases as everybody else's -- a quarter of all keys falls into ONE hash partition (ADVICE r1 / VERDICT r1 weak #8)."""
    rng = np.random.default_rng(seed)
    primer = "GATTACAGGCCTTAGCAATC"
    g = _rand_seq(rng, 60000)
    reads = []
    for _ in range(n_reads):
        st = int(rng.integers(0, len(g) - 60))
        s = primer + g[st:st + 60]
        reads.append(s if rng.integers(0, 2) else rc(s))
    reads += _sample(rng, g[:20000], 6000, 80, 80)             # plain reads of the same genome: overlaps among themselves and into the amplicons
    return from_strings(reads, min_overlap, "primer_prefixed")


def paired_mixed(seed=18, min_overlap=30):
    """Paired reads of mixed lengths at high coverage: most reads are contained in longer ones, so the mate-pair pass has to
    redirect mates to super reads and decide their orientation by substring search (Dataset.cpp:280-292)."""
    g = [synth.genome(seed, 5000)]
    b, o = synth.sample_reads(seed, g, 3000, 50, 110, paired=True, insert_mean=260.0, insert_sd=25.0)
    return dict(name="paired_mixed", bases=b, offsets=o, min_overlap=min_overlap, paired=True)


def small_configs():
    return [
        synth.config(1, scale=0.3),
        synth.config(2, scale=0.004),
        synth.config(3, scale=0.0015),
        synth.config(4, scale=0.0003),
        synth.containment_stress(5, genome_len=20000, n_primary=5000),
    ]


def adversarial():
    return [repeats(), repeats(mixed=True), tandem(), tandem(mixed=True), palindromes(), even_h(), filtered(), one_window()]
