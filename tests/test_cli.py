"""The command-line tool: portello's flags, messages and exit codes (src/cli.rs:8-170,
src/main.rs:24-122) on CPU; on a GPU the whole file-to-file run: indexed BAMs + FASTA in, remapped / unassembled BAMs out,
every output record byte-identical to the in-memory path (which the parity suite ties to the oracle)."""
import os
import subprocess

import numpy as np
import pytest

import helpers
from portello_b200 import abi, bamio, lib, synth
from test_bam_io import python_bam_records

CLI = os.path.join(lib.CSRC, "portello-b200")


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    lib.load()  # builds the library and the tool if missing
    s = synth.make("tiny", seed=23, n_reads=3000)
    paths = bamio.write_dataset(s, str(tmp_path_factory.mktemp("cli")), n_unmapped=40)
    return s, paths


def run(args, **kw):
    return subprocess.run([CLI] + args, capture_output=True, **kw)


def std_args(paths, out, un, extra=()):
    return ["--assembly-to-ref", paths["contigs"], "--read-to-assembly", paths["reads"], "--ref", paths["ref"], "--remapped-read-output", out,
            "--unassembled-read-output", un, "--threads", "4", *extra]


def test_version_help_and_missing_arguments():
    r = run(["--version"])
    assert r.returncode == 0 and r.stdout.decode().startswith("portello-b200 ")
    r = run(["--help"])
    assert r.returncode == 0 and b"--assembly-to-ref" in r.stdout and b"--unassembled-read-output" in r.stdout
    r = run(["--ref", "x.fa"])
    assert r.returncode == 2 and b"required arguments were not provided" in r.stderr
    r = run(["--bogus"])
    assert r.returncode == 2 and b"unexpected argument" in r.stderr


def test_usage_errors_match_the_reference_messages(dataset, tmp_path):
    _, paths = dataset
    out, un = str(tmp_path / "o.bam"), str(tmp_path / "u.bam")
    a = std_args(paths, out, un)
    bad = list(a)
    bad[1] = str(tmp_path / "missing.bam")
    r = run(bad)
    assert r.returncode == 64 and b"Invalid command-line setting: Can't find specified contig-to-ref bam file" in r.stderr  # cli.rs:97-100,137-140
    bad = list(a)
    bad[7] = str(tmp_path / "no_such_dir" / "o.bam")
    r = run(bad)
    assert r.returncode == 64 and b"Can't find existing directory for remapped read output file" in r.stderr              # cli.rs:110-117
    r = run(a[:-2] + ["--threads", "0"])
    assert r.returncode == 64 and b"--threads argument must be greater than 0" in r.stderr                               # cli.rs:125-127
    # an input without index: the reference's IndexedReader::from_path fails (cli.rs:149-154)
    noidx = tmp_path / "noindex.bam"
    noidx.write_bytes(open(paths["reads"], "rb").read())
    bad = list(a)
    bad[3] = str(noidx)
    r = run(bad)
    assert r.returncode == 101 and b"Failed to open input alignment file" in r.stderr


def test_reference_consistency_errors_exit_dataerr(dataset, tmp_path):
    """get_chrom_array (main.rs:24-62): a chromosome missing from the FASTA or of another length -> logged errors, exit 65."""
    s, paths = dataset
    fa = tmp_path / "short.fa"
    seqs = s.reference_arrays()
    bamio.write_fasta(str(fa), s.chrom_names, [seqs[0][:-7]] + list(seqs[1:-1]))  # first one short, last one missing
    a = std_args(paths, str(tmp_path / "o.bam"), str(tmp_path / "u.bam"))
    a[5] = str(fa)
    r = run(a)
    err = r.stderr.decode()
    assert r.returncode == 65, err
    assert "specified with inconsistent length" in err and "but not in the reference fasta" in err
    assert "[ERROR] Exiting due to one or more reference consistency issues" in err
    assert "][portello-b200][INFO] Starting portello-b200" in err  # fern line format [date][time][name][LEVEL] (logger.rs:11-19)


def test_no_cpu_fallback(dataset, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _, paths = dataset
    r = run(std_args(paths, str(tmp_path / "o.bam"), str(tmp_path / "u.bam")))
    assert r.returncode == 101 and b"no CPU fallback" in r.stderr


def expected_records(s, paths):
    """The in-memory path on the decoded file: every output record as bytes, per read."""
    f = bamio.BamFile(paths["reads"])
    dec = f.fetch(bamio.FETCH_ALL, flt=bamio.SKIP_SUPPLEMENTARY | bamio.SKIP_UNMAPPED_SECONDARY)
    a = dec.arrays()
    gctx = helpers.gpu_context(s)
    gctx.set_names(s.contig_names, s.chrom_names)
    pb = lib.PackedBatch(lib.load(), dec.recs, 0, dec.n, s.contig_names)
    res = helpers.lift_c(gctx, pb.c)
    x = dict(name_off=a["name_off"], names=np.concatenate([a["names"], np.zeros(32, np.uint8)]), aux_off=a["aux_off"],
             aux=np.concatenate([a["aux"], np.zeros(32, np.uint8)]), mate_tid=a["mate_tid"], mate_pos=a["mate_pos"], tlen=a["tlen"],
             qual=np.concatenate([a["qual"], np.zeros(32, np.uint8)]), qual_off=a["qual_off"])
    _, (rb, by) = gctx.assemble_records(x)
    raw = by.tobytes()
    octx = helpers.oracle_context(s)
    octx.set_names(s.contig_names, s.chrom_names)
    assert helpers.lift_c(octx, pb.c).diff(res) is None
    _, (rbo, byo) = octx.assemble_records(x)
    assert np.array_equal(rbo, rb) and np.array_equal(byo, by)  # the in-memory records are the oracle's
    return sorted(raw[int(rb[k]): int(rb[k + 1])] for k in range(len(rb) - 1)), res


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["file", "stdout"])
def test_file_to_file_run_equals_in_memory_path(dataset, tmp_path, mode):
    s, paths = dataset
    out, un = str(tmp_path / "remapped.bam"), str(tmp_path / "unassembled.bam")
    if mode == "file":
        r = run(std_args(paths, out, un, ("--batch-reads", "700")))
    else:  # the reference's pipe mode: uncompressed BAM on stdout (read_alignment_scanner.rs:66-71)
        r = run(std_args(paths, "-", un))
        open(out, "wb").write(r.stdout)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert b"completed. Total Runtime:" in r.stderr
    text, refs, recs = python_bam_records(out)
    assert [n for n, _ in refs] == s.chrom_names and [ln for _, ln in refs] == [int(s.chrom_len[i]) for i in range(s.n_chrom)]
    assert text.startswith("@HD\tVN:1.6\tSO:unsorted\n") and "@PG\tPN:portello-b200" in text and text.count("@SQ") == s.n_chrom
    want, res = expected_records(s, paths)
    got = sorted(p["raw"] for p in recs)
    assert len(got) == len(want) == res.n_records
    assert got == want
    # records of one read stay together and in order (written under one lock, :482-487)
    names = [p["name"] for p in recs]
    seen, last = set(), None
    for nm in names:
        if nm != last:
            assert nm not in seen, "records of a read are interleaved with another read's"
            seen.add(nm)
            last = nm
    # unmapped reads pass through unchanged (:537-559)
    _, refs_u, recs_u = python_bam_records(un)
    _, _, py_in = python_bam_records(paths["reads"])
    assert [p["raw"] for p in recs_u] == [p["raw"] for p in py_in if p["flag"] & 4] and len(recs_u) == 40
    assert [n for n, _ in refs_u] == s.chrom_names
    if mode == "stdout":  # level 0: every BGZF block of the record stream is a stored block
        raw = open(out, "rb").read()
        assert len(raw) > sum(len(p["raw"]) for p in recs)
    # the output is itself a valid input of the reader (round trip of the output half through the input half)
    f = bamio.BamFile(out)
    assert f.has_eof_marker and f.fetch(bamio.FETCH_ALL).n == len(recs)
