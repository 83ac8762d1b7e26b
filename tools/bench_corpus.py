"""BASELINE config 5 at reduced scale: decode-only of a corpus of short reference-format files with mixed
presets (block sizes 4096 / 12288 / 16384).  Streams are produced by this library's encoder (byte-identical
to the reference's, see tests); decode is timed file by file (SLADecoder_DecodeWhole), through the batch
entry point (SLAB200_Decoder_DecodeBatchPCM) and with the reference on all host cores.
usage: python tools/bench_corpus.py [num_files] > corpus.json"""
import ctypes as C, json, multiprocessing as mp, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from sla_b200 import capi, synth

NFILES = int(sys.argv[1]) if len(sys.argv) > 1 else 512
RATE, BITS, NCH = 44100, 16, 2
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libsla_ref.so")


def make_corpus(lib):
    rng = np.random.default_rng(0xC5)
    bases = [synth.synth_pcm(NCH, 30 * RATE, BITS, RATE, 500 + k) for k in range(6)]
    files = []
    for i in range(NFILES):
        n = int(rng.integers(3 * RATE, 30 * RATE + 1))
        base = bases[i % len(bases)]
        start = int(rng.integers(0, base.shape[1] - n + 1))
        pcm = np.ascontiguousarray(base[:, start:start + n] if (i // len(bases)) % 2 == 0 else -np.maximum(base[:, start:start + n], -(2 ** 31 - 65536)))
        preset = (0, 2, 4)[int(rng.integers(0, 3))]
        rc, data = lib.encode_whole(pcm, BITS, RATE, capi.preset_parameter(preset, NCH))
        assert rc == 0
        files.append((data, pcm, preset))
    return files


def _ref_decode(data):
    lib = capi.SLALibrary(REF_SO)
    t0 = time.perf_counter()
    rc, pcm, _ = lib.decode_whole(data)
    return time.perf_counter() - t0, rc, pcm.size


def _ref_encode(args):
    pcm, preset = args
    lib = capi.SLALibrary(REF_SO)
    t0 = time.perf_counter()
    rc, data = lib.encode_whole(pcm, BITS, RATE, capi.preset_parameter(preset, NCH))
    return time.perf_counter() - t0, rc, pcm.size


def bench_encode(lib, files, torch):
    """Encode leg: the same corpus, PCM in pinned host memory, one call per file against the batch call."""
    L = lib.lib
    L.SLAB200_Encoder_EncodeBatchPCM.argtypes = [C.c_void_p, C.POINTER(capi.EncodeItem), C.c_uint32]
    L.SLAB200_Encoder_EncodePCM.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    fb = NCH * BITS // 8
    ioff, itotal = [], 0
    for _, pcm, _ in files:
        ioff.append(itotal); itotal += (pcm.shape[1] * fb + 63) & ~63
    h_pcm = torch.empty(itotal, dtype=torch.uint8, pin_memory=True)
    for (_, pcm, _), o in zip(files, ioff):
        h_pcm.numpy()[o:o + pcm.shape[1] * fb] = np.frombuffer(capi.planar_to_pcm(pcm, BITS), dtype=np.uint8)
    ooff, ototal = [], 0
    for d, _, _ in files:
        ooff.append(ototal); ototal += (2 * len(d) + 65536 + 63) & ~63
    h_sla = torch.empty(ototal, dtype=torch.uint8, pin_memory=True)
    enc = L.SLAEncoder_Create(C.byref(capi.EncoderConfig(**capi.CLI_CAPACITY, verpose_flag=0)))
    wf = capi.WaveFormat(NCH, BITS, RATE, 0)
    by_preset = {}
    for i, (_, _, preset) in enumerate(files):
        by_preset.setdefault(preset, []).append(i)

    def setup(preset):
        ep = capi.preset_parameter(preset, NCH)
        assert L.SLAEncoder_SetWaveFormat(enc, C.byref(wf)) == 0 and L.SLAEncoder_SetEncodeParameter(enc, C.byref(ep)) == 0

    def check():
        return all(bytes(h_sla.numpy()[ooff[i]:ooff[i] + len(files[i][0])]) == files[i][0] for i in range(len(files)))

    def run_file_by_file():
        size = C.c_uint32(0)
        for preset, idx in by_preset.items():
            setup(preset)
            for i in idx:
                rc = L.SLAB200_Encoder_EncodePCM(enc, h_pcm.data_ptr() + ioff[i], files[i][1].shape[1], h_sla.data_ptr() + ooff[i],
                                                 2 * len(files[i][0]) + 65536, C.byref(size))
                assert rc == 0 and size.value == len(files[i][0])

    def run_batch():
        for preset, idx in by_preset.items():
            setup(preset)
            items = (capi.EncodeItem * len(idx))()
            for k, i in enumerate(idx):
                items[k].pcm = h_pcm.data_ptr() + ioff[i]; items[k].num_samples = files[i][1].shape[1]
                items[k].data = h_sla.data_ptr() + ooff[i]; items[k].data_size = 2 * len(files[i][0]) + 65536
            rc = L.SLAB200_Encoder_EncodeBatchPCM(enc, items, len(idx))
            assert rc == 0 and all(items[k].result == 0 and items[k].output_size == len(files[idx[k]][0]) for k in range(len(idx)))

    run_file_by_file(); h_sla.zero_()
    t0 = time.perf_counter(); run_file_by_file(); t_file = time.perf_counter() - t0
    ok_a = check()
    run_batch(); h_sla.zero_()
    t0 = time.perf_counter(); run_batch(); t_batch = time.perf_counter() - t0
    ok_b = check()
    L.SLAEncoder_Destroy(enc)
    cpu = None
    if os.path.exists(REF_SO):
        cores = os.cpu_count() or 1
        sample = [(p, pr) for _, p, pr in files[:min(len(files), 2 * cores)]]
        with mp.get_context("fork").Pool(cores) as pool:
            r = pool.map(_ref_encode, sample)
        busy = sum(x[0] for x in r)
        cpu = {"value": sum(x[2] for x in r) / (busy / cores) / 1e6, "cores": cores, "files": len(sample), "kind": "reference",
               "per_core": sum(x[2] for x in r) / busy / 1e6}
    chsamp = sum(p.size for _, p, _ in files)
    return {"file_by_file": {"value": chsamp / t_file / 1e6, "seconds": t_file, "byte_identical": ok_a, "api": "SLAB200_Encoder_EncodePCM per file, one handle"},
            "batch": {"value": chsamp / t_batch / 1e6, "seconds": t_batch, "byte_identical": ok_b,
                      "api": "SLAB200_Encoder_EncodeBatchPCM per preset, pinned host buffers, H2D + kernels + D2H inside",
                      "workers": int(os.environ.get("SLAB200_BATCH_ENC_WORKERS", "8"))},
            "cpu_baseline": cpu}


def main():
    real_stdout = os.dup(1); os.dup2(2, 1)
    lib = capi.SLALibrary(os.path.join(ROOT, "sla_b200", "lib", "libsla_b200.so"))
    files = make_corpus(lib)
    chsamp = sum(p.size for _, p, _ in files)
    streams = [d for d, _, _ in files]
    import torch
    if os.environ.get("CORPUS_ENCODE_ONLY"):
        os.write(real_stdout, (json.dumps({"encode": bench_encode(lib, files, torch)}) + "\n").encode())
        return
    L = lib.lib
    L.SLAB200_Decoder_DecodeBatchPCM.argtypes = [C.c_void_p, C.POINTER(capi.BatchItem), C.c_uint32]
    L.SLAB200_Decoder_DecodePCM.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    dec = L.SLADecoder_Create(C.byref(capi.DecoderConfig(**capi.CLI_CAPACITY, enable_crc_check=1, verpose_flag=0)))
    fb = NCH * BITS // 8
    # pinned host buffers: all streams back to back, all PCM outputs back to back
    offs, total = [], 0
    for d in streams:
        offs.append(total); total += (len(d) + 63) & ~63
    h_in = torch.empty(total, dtype=torch.uint8, pin_memory=True)
    for d, o in zip(streams, offs):
        h_in.numpy()[o:o + len(d)] = np.frombuffer(d, dtype=np.uint8)
    outs, ototal = [], 0
    for _, pcm, _ in files:
        outs.append(ototal); ototal += pcm.shape[1] * fb
    h_out = torch.empty(ototal, dtype=torch.uint8, pin_memory=True)
    want = np.concatenate([np.frombuffer(capi.planar_to_pcm(pcm, BITS), dtype=np.uint8) for _, pcm, _ in files])

    def run_file_by_file():
        got = C.c_uint32(0)
        for (d, pcm, _), o, oo in zip(files, offs, outs):
            rc = L.SLAB200_Decoder_DecodePCM(dec, h_in.data_ptr() + o, len(d), h_out.data_ptr() + oo, pcm.shape[1], C.byref(got))
            assert rc == 0 and got.value == pcm.shape[1]

    def run_batch():
        items = (capi.BatchItem * len(files))()
        for i, ((d, pcm, _), o, oo) in enumerate(zip(files, offs, outs)):
            items[i].data = h_in.data_ptr() + o; items[i].data_size = len(d)
            items[i].pcm = h_out.data_ptr() + oo; items[i].capacity_samples = pcm.shape[1]
        rc = L.SLAB200_Decoder_DecodeBatchPCM(dec, items, len(files))
        assert rc == 0 and all(items[i].result == 0 for i in range(len(files)))

    # (a) one call per file, one handle (SLAB200_Decoder_DecodePCM)
    run_file_by_file()
    h_out.zero_()
    t0 = time.perf_counter(); run_file_by_file(); t_file = time.perf_counter() - t0
    ok_a = bool(np.array_equal(h_out.numpy(), want))
    # (b) one call for the corpus (SLAB200_Decoder_DecodeBatchPCM)
    run_batch()
    h_out.zero_()
    t0 = time.perf_counter(); run_batch(); t_batch = time.perf_counter() - t0
    ok_b = bool(np.array_equal(h_out.numpy(), want))
    # (c) the reference on all host cores, a bounded sample
    cpu = None
    if os.path.exists(REF_SO):
        cores = os.cpu_count() or 1
        sample = streams[:min(len(streams), 8 * cores)]
        with mp.get_context("fork").Pool(cores) as pool:
            r = pool.map(_ref_decode, sample)
        busy = sum(x[0] for x in r)                  # decode time only; the cores run concurrently
        cpu = {"value": sum(x[2] for x in r) / (busy / cores) / 1e6, "cores": cores, "files": len(sample), "kind": "reference",
               "per_core": sum(x[2] for x in r) / busy / 1e6}
    line = {"config": "C5 (reduced): %d stereo 16-bit 44.1 kHz files of 3-30 s, presets {0,2,4} mixed" % NFILES,
            "channel_samples": chsamp, "unit": "M channel-samples/s",
            "file_by_file": {"value": chsamp / t_file / 1e6, "seconds": t_file, "bit_exact": ok_a, "api": "SLAB200_Decoder_DecodePCM per file, one handle"},
            "batch": {"value": chsamp / t_batch / 1e6, "seconds": t_batch, "bit_exact": ok_b,
                      "api": "SLAB200_Decoder_DecodeBatchPCM, pinned host buffers, H2D + kernels + D2H inside"},
            "cpu_baseline": cpu}
    line["encode"] = bench_encode(lib, files, torch)
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()
