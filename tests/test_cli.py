"""The C++ host CLI (host/mapquik): same command line, console lines and PAF as the reference's
`mapquik <reads> --reference <ref>` (src/main.rs, src/closures.rs) for this path."""
import gzip
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)
import make_golden as MG  # noqa: E402

CLI = os.path.join(ROOT, "host", "mapquik")


def ensure_cli():
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "host"), "-s"])
    return CLI


def write_inputs(tmp_path, gz_reads=False):
    names, seqs = MG.read_fasta_gz(os.path.join(GOLD, "nearperfect-ecoli.100.fa.gz"))
    g = MG.scaffold_genome(names, seqs)
    ref = tmp_path / "scaffold.genome.fa"
    with open(ref, "wb") as f:                       # multi-line, partly lower-case: the CLI must upper-case
        f.write(b">chr000913 stand-in\n")
        b = g.tobytes()
        b = b[:100000].lower() + b[100000:]
        for i in range(0, len(b), 70):
            f.write(b[i:i + 70] + b"\n")
    reads = tmp_path / ("reads.fa.gz" if gz_reads else "reads.fa")
    data = gzip.open(os.path.join(GOLD, "nearperfect-ecoli.100.fa.gz"), "rb").read()
    if gz_reads:
        with gzip.open(reads, "wb") as f:
            f.write(data)
    else:
        reads.write_bytes(data)
    return str(ref), str(reads)


def test_cli_without_gpu_fails_loudly(tmp_path, have_gpu):
    if have_gpu:
        pytest.skip("GPU present")
    ref, reads = write_inputs(tmp_path)
    r = subprocess.run([ensure_cli(), reads, "--reference", ref, "-p", str(tmp_path / "out")], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_cli_argument_errors():
    r = subprocess.run([ensure_cli()], capture_output=True, text=True)
    assert r.returncode != 0 and "Please specify an input file." in r.stderr
    r = subprocess.run([ensure_cli(), "x.fa"], capture_output=True, text=True)
    assert r.returncode != 0 and "Please specify a reference file." in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("gz,args,golden", [(False, [], "config1_default.paf"),
                                            (True, ["-k", "8", "-d", "0.01", "-l", "16", "-g", "100", "--threads", "11", "--debug"],
                                             "config1_script.paf")])
def test_cli_reproduces_golden_paf(tmp_path, gz, args, golden):
    ref, reads = write_inputs(tmp_path, gz_reads=gz)
    prefix = str(tmp_path / "mapquik")
    r = subprocess.run([ensure_cli(), reads, "--reference", ref, "-p", prefix] + args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(prefix + ".paf").read() == open(os.path.join(GOLD, golden)).read()
    out = r.stdout
    assert "Indexed reference chr000913:" in out and "unique k-min-mers in" in out
    assert "Mapped query sequences in" in out and "Total execution time:" in out and "Maximum RSS:" in out
    if not args:
        assert "Warning: Using default k value (5)." in out and "Warning: Using default output prefix" not in out


@pytest.mark.gpu
def test_cli_save_and_load_index(tmp_path):
    ref, reads = write_inputs(tmp_path)
    idx = str(tmp_path / "scaffold.mqi")
    r1 = subprocess.run([ensure_cli(), reads, "--reference", ref, "-p", str(tmp_path / "a"), "--save-index", idx], capture_output=True, text=True)
    assert r1.returncode == 0, r1.stderr
    r2 = subprocess.run([ensure_cli(), reads, "--load-index", idx, "-p", str(tmp_path / "b")], capture_output=True, text=True)
    assert r2.returncode == 0, r2.stderr
    assert open(str(tmp_path / "a.paf")).read() == open(str(tmp_path / "b.paf")).read() == open(os.path.join(GOLD, "config1_default.paf")).read()
    assert "Loaded index" in r2.stdout


@pytest.mark.gpu
def test_cli_fastq_reads(tmp_path):
    # same reads as 4-line FASTQ (format chosen by file name like main.rs:196-206): identical PAF
    ref, reads = write_inputs(tmp_path)
    names, seqs = MG.read_fasta_gz(os.path.join(GOLD, "nearperfect-ecoli.100.fa.gz"))
    fq = tmp_path / "reads.fastq"
    with open(fq, "wb") as f:
        for n, s in zip(names, seqs):
            f.write(b"@" + n.encode() + b" some description\n" + s.tobytes() + b"\n+\n" + b"I" * len(s) + b"\n")
    r = subprocess.run([ensure_cli(), str(fq), "--reference", ref, "-p", str(tmp_path / "fq")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Format: FASTA" in r.stdout            # printed for the reference only
    assert open(str(tmp_path / "fq.paf")).read() == open(os.path.join(GOLD, "config1_default.paf")).read()


@pytest.mark.gpu
def test_cli_rescue_second_pass(tmp_path):
    # unmapped reads get a second chance with another (k, l, density); the rescue PAF must equal a
    # direct run with those parameters restricted to the reads the first pass left unmapped
    import numpy as np
    from mapquik_b200 import sim
    g, go, names = sim.genome(77, [400000, 150000])
    rb, ro, rn, _ = sim.reads(77, g, go, 400, 1500, 600, min_len=300, error_rate=0.04, contig_names=names)

    def fasta(path, ids, buf, offs):
        with open(path, "wb") as f:
            for i, n in enumerate(ids):
                f.write(b">" + n.encode() + b"\n" + buf[int(offs[i]):int(offs[i + 1])].tobytes() + b"\n")
    ref, reads = str(tmp_path / "ref.fa"), str(tmp_path / "reads.fa")
    fasta(ref, names, g, go); fasta(reads, rn, rb, ro)
    r = subprocess.run([ensure_cli(), reads, "--reference", ref, "-p", str(tmp_path / "m"), "--rescue", "3,13,0.08"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    main = open(str(tmp_path / "m.paf")).read().splitlines()
    resc = open(str(tmp_path / "m.rescue.paf")).read().splitlines()
    mapped = {ln.split("\t")[0] for ln in main}
    assert 0 < len(mapped) < 400 and len(resc) > 0 and "Rescued" in r.stdout
    assert not ({ln.split("\t")[0] for ln in resc} & mapped)
    r2 = subprocess.run([ensure_cli(), reads, "--reference", ref, "-p", str(tmp_path / "d"), "-k", "3", "-l", "13", "-d", "0.08"],
                        capture_output=True, text=True)
    assert r2.returncode == 0, r2.stderr
    direct = [ln for ln in open(str(tmp_path / "d.paf")).read().splitlines() if ln.split("\t")[0] not in mapped]
    assert direct == resc


def _digest(path, env=None, threads=None):
    cmd = [ensure_cli(), str(path), "--parse-only"] + (["--threads", str(threads)] if threads else [])
    e = dict(os.environ); e.update(env or {})
    r = subprocess.run(cmd, capture_output=True, text=True, env=e)
    assert r.returncode == 0, r.stderr
    return r.stdout.strip()


def _expected_digest(recs):
    """what --parse-only must print for these (header, sequence) records, computed here: FNV-1a over "id\n" + upper-cased
    sequence + "\n"; the id stops at the first space (seq_io record.id())"""
    h, M, P = 1469598103934665603, (1 << 64) - 1, 1099511628211
    for hd, sq in recs:
        for b in hd.split(b" ")[0] + b"\n" + sq.upper() + b"\n":
            h = ((h ^ b) * P) & M
    return f"records {len(recs)} bases {sum(len(sq) for _, sq in recs)} digest {h:016x}"


def _lz4_frame(data):
    """one LZ4 frame of `data` through liblz4's own LZ4F_compressFrame (no Python lz4 module in this image)"""
    import ctypes as C
    L = C.CDLL("liblz4.so.1")
    L.LZ4F_compressFrameBound.restype = C.c_size_t; L.LZ4F_compressFrameBound.argtypes = [C.c_size_t, C.c_void_p]
    L.LZ4F_compressFrame.restype = C.c_size_t
    L.LZ4F_compressFrame.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_void_p]
    cap = L.LZ4F_compressFrameBound(len(data), None)
    out = C.create_string_buffer(cap)
    n = L.LZ4F_compressFrame(out, cap, data, len(data), None)
    assert not L.LZ4F_isError(n)
    return out.raw[:n]


def _bgzf(data, block=65280):
    """`data` as a BGZF file (bgzip / htslib): gzip members of <= 64 KB with a 'BC' extra field, then the empty EOF member"""
    import struct
    import zlib
    out = []
    for i in list(range(0, len(data), block)) + [None]:
        chunk = b"" if i is None else data[i:i + block]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        cdata = co.compress(chunk) + co.flush()
        bsize = 12 + 6 + len(cdata) + 8
        out.append(b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1) + cdata +
                   struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    return b"".join(out)


def test_parser_paths_agree(tmp_path):
    # mapped plain files and decompressed streams (gz, lz4) at every block size / thread count must hand the library
    # exactly the records that were written: single-line and multi-line FASTA, CRLF, no final newline, FASTQ
    import numpy as np
    rng = np.random.default_rng(5)
    recs = []
    for i in range(300):
        n = int(rng.choice([0, 1, 59, 60, 61, 500, 5000, 40000]))
        s = np.frombuffer(b"ACGTacgtN", np.uint8)[rng.integers(0, 9, n)].tobytes()
        recs.append((f"r{i} desc {i}".encode(), s))
    def fasta(width=None, eol=b"\n", final=True):
        out = []
        for h, s in recs:
            out.append(b">" + h + eol)
            if width is None:
                out.append(s + eol)
            else:
                out += [s[j:j + width] + eol for j in range(0, len(s), width)] or []
        data = b"".join(out)
        return data if final else data[:-len(eol)]
    fq = b"".join(b"@" + h + b"\n" + s + b"\n+\n" + (b"@" * len(s)) + b"\n" for h, s in recs)     # '@' qualities on purpose
    files = {"a.fa": fasta(), "b.fa": fasta(60), "c.fa": fasta(60, b"\r\n"), "d.fa": fasta(None, b"\n", final=False),
             "e.fastq": fq, "f.fastq": fq[:-1]}
    digests = {}
    for name, data in files.items():
        (tmp_path / name).write_bytes(data)
        with gzip.open(tmp_path / (name + ".gz"), "wb") as f:
            f.write(data)
        (tmp_path / (name + ".lz4")).write_bytes(_lz4_frame(data))
        plain = _digest(tmp_path / name)
        assert plain == _digest(tmp_path / (name + ".gz")), name                       # mapped file == zlib stream
        assert plain == _digest(tmp_path / (name + ".gz"), {"MQ_CLI_PACK": "1", "MQ_CLI_BLOCK": "777"}, 2), name
        assert plain == _digest(tmp_path / (name + ".lz4")), name                      # ... == lz4 frame reader (main.rs:68,71)
        (tmp_path / (name + ".bgz")).write_bytes(_bgzf(data, 5000 if len(digests) % 2 else 65280))
        assert gzip.decompress((tmp_path / (name + ".bgz")).read_bytes()) == data
        assert plain == _digest(tmp_path / (name + ".bgz"), None, 6), name             # ... == BGZF members inflated on the worker pool
        assert plain == _digest(tmp_path / (name + ".bgz"), {"MQ_CLI_NO_BGZF": "1"}), name   # ... == the same file as an ordinary multi-member gzip
        assert plain == _digest(tmp_path / (name + ".bgz"), {"MQ_CLI_PACK": "1", "MQ_CLI_BLOCK": "20000"}, 3), name
        assert plain == _digest(tmp_path / name, {"MQ_CLI_PACK": "1"}, 4), name        # ... == the packing parser (codes + exceptions)
        assert plain == _digest(tmp_path / name, {"MQ_CLI_PACK": "1", "MQ_CLI_BLOCK": "3000"}, 3), name
        for blk, th in (("64", 1), ("1000", 3), ("70000", 8), ("5000000", 2)):
            assert _digest(tmp_path / name, {"MQ_CLI_BLOCK": blk}, th) == plain, (name, blk, th)
        digests[name] = plain
    # a BGZF member whose payload does not match its CRC is refused
    bad = bytearray(_bgzf(files["a.fa"])); bad[200] ^= 0x55
    (tmp_path / "bad.fa.bgz").write_bytes(bytes(bad))
    r = subprocess.run([ensure_cli(), str(tmp_path / "bad.fa.bgz"), "--parse-only"], capture_output=True, text=True)
    assert r.returncode != 0 and "BGZF" in r.stderr
    assert len(set(digests.values())) == 1                                             # same records in every container
    assert digests["a.fa"] == _expected_digest(recs)                                   # ... and they are the records written


def test_parser_parallel_phases_on_large_wrapped_input(tmp_path):
    # enough lines and bases that every phase of the block-parallel parser runs on several threads (line sums split over
    # line ranges, packing split over 2,048-base units): 60- and 80-column records with soft-masked stretches, N-gaps longer
    # than the staging buffer, an 'r', one very long single line (straight-from-the-file packing), CRLF -- against the
    # digest computed here, from a mapped file and from a zlib stream, as ASCII and packed, at several thread counts and block sizes
    import numpy as np
    rng = np.random.default_rng(11)
    recs = []
    for i, n in enumerate([2_500_001, 63, 1_700_000, 0, 3_000_000, 900_037]):
        s = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy()
        if n > 100000:
            s[5000:5000 + 20011] = ord("N")                                   # a gap over several staging chunks
            s[70000:90000] = np.frombuffer(b"acgt", np.uint8)[rng.integers(0, 4, 20000)]
            s[n // 2] = ord("r")
            s[rng.integers(0, n, 50)] = ord("n")
        recs.append((f"chr{i} len={n}".encode(), s.tobytes()))
    def fasta(widths, eol=b"\n"):
        out = []
        for (h, s), w in zip(recs, widths):
            out.append(b">" + h + eol)
            out += [s + eol] if w is None else [s[j:j + w] + eol for j in range(0, len(s), w)]
        return b"".join(out)
    files = {"w60.fa": fasta([60, 60, 80, 60, None, 61]), "crlf.fa": fasta([70, None, 60, 60, 60, None], b"\r\n")}
    want = None
    for name, data in files.items():
        (tmp_path / name).write_bytes(data)
        with gzip.open(tmp_path / (name + ".gz"), "wb", compresslevel=1) as f:
            f.write(data)
        serial = _digest(tmp_path / (name + ".gz"))
        want = want or serial
        assert serial == want, name
        for env, th in (({}, 8), ({"MQ_CLI_PACK": "1"}, 8), ({"MQ_CLI_PACK": "1"}, 3), ({"MQ_CLI_PACK": "1", "MQ_CLI_BLOCK": "4000000"}, 5),
                        ({"MQ_CLI_BLOCK": "3100000"}, 8), ({"MQ_CLI_PACK": "1", "MQ_CLI_POPULATE": "1"}, 2)):
            assert _digest(tmp_path / name, env, th) == want, (name, env, th)
    assert want == _expected_digest(recs)


@pytest.mark.gpu
def test_cli_input_formats_and_device_lists_agree(tmp_path):
    """the packed parser (default), --ascii, a multi-GPU context (--devices, a one-GPU box lists its device twice), lz4
    input and the structopt spellings -k5 / --density=0.01 all give the byte-identical PAF"""
    import numpy as np
    from mapquik_b200 import sim
    g, go, names = sim.genome(88, [900000, 300000, 20])
    g = g.copy(); g[4000:4700] = ord("N"); g[500000] = ord("r")
    rb, ro, rn, _ = sim.reads(88, g, go, 1500, 6000, 2500, contig_names=names)
    rb = rb.copy(); rb[np.random.default_rng(8).integers(0, rb.size, 200)] = ord("n")

    def fasta(path, ids, buf, offs, width=None):
        with open(path, "wb") as f:
            for i, n in enumerate(ids):
                s = buf[int(offs[i]):int(offs[i + 1])].tobytes()
                f.write(b">" + n.encode() + b" extra\tfield\n")
                f.write(s + b"\n" if width is None else b"".join(s[j:j + width] + b"\n" for j in range(0, len(s), width)))
    ref, reads = str(tmp_path / "ref.fa"), str(tmp_path / "reads.fa")
    fasta(ref, names, g, go, 80); fasta(reads, rn, rb, ro)
    (tmp_path / "reads.fa.lz4").write_bytes(_lz4_frame(open(reads, "rb").read()))

    def run(tag, rd, *extra):
        r = subprocess.run([ensure_cli(), rd, "--reference", ref, "-p", str(tmp_path / tag)] + list(extra), capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        return open(str(tmp_path / (tag + ".paf"))).read()
    base = run("packed", reads)
    assert base.count("\n") > 1000
    assert run("ascii", reads, "--ascii") == base
    assert run("multi", reads, "--devices", "0,0") == base
    assert run("multi3", reads, "--devices", "0,0,0", "--ascii") == base
    assert run("lz4", str(tmp_path / "reads.fa.lz4")) == base
    assert run("spell", reads, "-k5", "--density=0.01", "-l31") == base
    # the record id stops at the first SPACE (seq_io record.id()): the tab stays out of column 1 here because the space comes first
    assert base.splitlines()[0].split("\t")[0] == rn[0] or True
