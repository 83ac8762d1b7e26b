"""The command-line tool (sla_b200/cli/sla_b200_cli.c) against the reference tool
(oracle/_ref/sla_ref_cli = unmodified src/main.c + src/wav.c): same options, byte-identical .sla and .wav
files, same exit codes.  On CPU the tool runs on the host-simulator build of the kernels; the GPU tests run
the product binary."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT

REF_CLI = os.path.join(ROOT, "oracle", "_ref", "sla_ref_cli")
HOSTSIM_CLI = os.path.join(ROOT, "tests", "hostsim", "sla_hostsim_cli")
PRODUCT_CLI = os.path.join(ROOT, "sla_b200", "lib", "sla_b200_cli")


def make_wav(path, nch, bits, rate, n, seed, extra_chunks=False):
    """A tone + noise file; with extra_chunks an 18-byte fmt chunk and a LIST chunk precede the data."""
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    full = 2.0 ** (bits - 1) - 1
    chans = []
    for c in range(nch):
        x = 0.4 * np.sin(t * (0.02 + 0.003 * c)) + 0.1 * np.sin(t * 0.11 + c) + rng.normal(0, 0.01, n)
        x[n // 3:n // 3 + 600] = 0.0
        chans.append(np.clip(np.round(x * full), -full - 1, full).astype(np.int64))
    frames = np.stack(chans, 1)
    if bits == 8:
        pcm = (frames + 128).astype(np.uint8).tobytes()
    elif bits == 16:
        pcm = frames.astype("<i2").tobytes()
    elif bits == 24:
        b = frames.astype("<i4").tobytes()
        pcm = np.frombuffer(b, dtype=np.uint8).reshape(-1, 4)[:, :3].tobytes()
    else:
        pcm = frames.astype("<i4").tobytes()
    fb = nch * bits // 8
    fmt = struct.pack("<HHIIHH", 1, nch, rate, rate * fb, fb, bits)
    if extra_chunks:
        fmt_chunk = b"fmt " + struct.pack("<I", 18) + fmt + b"\0\0"
        other = b"LIST" + struct.pack("<I", 10) + b"INFOabcdef"
    else:
        fmt_chunk = b"fmt " + struct.pack("<I", 16) + fmt
        other = b""
    body = b"WAVE" + fmt_chunk + other + b"data" + struct.pack("<I", len(pcm)) + pcm
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)


def run(cli, *args):
    return subprocess.run([cli, *map(str, args)], capture_output=True, text=True)


CASES = [  # nch, bits, rate, samples, preset, extra chunks
    (1, 8, 48000, 20000, 0, False),
    (2, 16, 44100, 30000, 2, True),
    (2, 24, 96000, 25000, 4, False),
    (3, 16, 32000, 9000, 3, False),
]


def _same_files_as_reference(cli, tmp_path, cases):
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/sla_ref_cli not built (needs /root/reference at build time)")
    assert os.path.exists(cli), f"{cli} missing: run `make cli`"
    for i, (nch, bits, rate, n, preset, extra) in enumerate(cases):
        wav = tmp_path / f"in{i}.wav"
        make_wav(wav, nch, bits, rate, n, seed=100 + i, extra_chunks=extra)
        ref_sla, our_sla = tmp_path / f"ref{i}.sla", tmp_path / f"our{i}.sla"
        assert run(REF_CLI, "-e", "-q", "-m", preset, wav, ref_sla).returncode == 0
        r = run(cli, "-e", "-q", "-m", preset, wav, our_sla)
        assert r.returncode == 0, r.stderr
        assert our_sla.read_bytes() == ref_sla.read_bytes()
        ref_wav, our_wav, strm_wav = tmp_path / f"ref{i}.wav", tmp_path / f"our{i}.wav", tmp_path / f"strm{i}.wav"
        assert run(REF_CLI, "-d", "-q", ref_sla, ref_wav).returncode == 0
        r = run(cli, "-d", "-q", ref_sla, our_wav)
        assert r.returncode == 0, r.stderr
        assert our_wav.read_bytes() == ref_wav.read_bytes()
        if bits <= 24:                                   # the reference CLI's streaming handle takes <= 24 bits
            r = run(cli, "-d", "-s", "-q", ref_sla, strm_wav)
            assert r.returncode == 0, r.stderr
            assert strm_wav.read_bytes() == ref_wav.read_bytes()
    # batch decode of everything at once
    out = tmp_path / "batch"
    out.mkdir()
    r = run(cli, "-d", "-q", "-b", out, *[tmp_path / f"ref{i}.sla" for i in range(len(cases))])
    assert r.returncode == 0, r.stderr
    for i in range(len(cases)):
        assert (out / f"ref{i}.wav").read_bytes() == (tmp_path / f"ref{i}.wav").read_bytes()
    # batch encode
    r = run(cli, "-e", "-q", "-m", cases[1][4], "-b", out, tmp_path / "in1.wav")
    assert r.returncode == 0, r.stderr
    assert (out / "in1.sla").read_bytes() == (tmp_path / "ref1.sla").read_bytes()


def test_cli_files_equal_reference_hostsim(tmp_path, hostsim):
    subprocess.run(["make", "-s", "-C", ROOT, HOSTSIM_CLI[len(ROOT) + 1:]], check=True)
    _same_files_as_reference(HOSTSIM_CLI, tmp_path, CASES[:3])


def test_cli_usage_and_errors_match_reference(tmp_path, hostsim):
    """Exit codes and messages of src/main.c:434-537 and src/command_line_parser.c."""
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/sla_ref_cli not built")
    subprocess.run(["make", "-s", "-C", ROOT, HOSTSIM_CLI[len(ROOT) + 1:]], check=True)
    wav = tmp_path / "x.wav"
    make_wav(wav, 1, 16, 8000, 3000, seed=1)
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"RIFF\x10\0\0\0WAVEjunkjunkjunkjunkjunkjunkjunkjunkjunk")
    name = lambda s, cli: s.replace(cli, "CLI")
    for args in ([], ["-h"], ["-v"], ["-e"], ["-e", wav], ["-e", "-d", wav, tmp_path / "o"], [wav, tmp_path / "o"],
                 ["-e", "-m", "7", wav, tmp_path / "o"], ["-e", "-m"], ["-x", wav, tmp_path / "o"], ["--nope"],
                 ["-e", "-e", wav, tmp_path / "o"], ["-em", wav, tmp_path / "o"], ["-e", bad, tmp_path / "o"],
                 ["-e", tmp_path / "missing.wav", tmp_path / "o"], ["-e", "--mode=1", "-q", wav, tmp_path / "o1"],
                 ["-d", "-q", "-c", "no", tmp_path / "o1", tmp_path / "o1.wav"]):
        a, b = run(REF_CLI, *args), run(HOSTSIM_CLI, *args)
        assert a.returncode == b.returncode, args
        if args != ["-h"]:                                # our help lists one more option (-b)
            assert name(a.stdout, REF_CLI) == name(b.stdout, HOSTSIM_CLI), args
        assert name(a.stderr, REF_CLI) == name(b.stderr, HOSTSIM_CLI), args
    h = run(HOSTSIM_CLI, "-h").stdout
    assert name(run(REF_CLI, "-h").stdout, REF_CLI).splitlines() == name(h, HOSTSIM_CLI).splitlines()[:-1]


@pytest.mark.gpu
def test_cli_files_equal_reference_gpu(tmp_path, product):
    _same_files_as_reference(PRODUCT_CLI, tmp_path, CASES)
