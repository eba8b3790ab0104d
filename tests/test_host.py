"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol the header
declares, the host marker parser agrees with the oracle on what can be decided without entropy decode,
the scheduler's partition rule, the Python mirror of the reference's types, and the N>1 plumbing of
bench.py under torch.distributed/gloo (world_size 2).  No compute call is made (no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O
from _damage import header_damage

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def zlib():
    from zpix_b200 import _lib

    return _lib


def _read(d, name):
    with open(os.path.join(d, name), "rb") as f:
        return f.read()


def test_library_exports_every_declared_symbol(zlib):
    hdr = open(os.path.join(ROOT, "include", "zpix_cuda.h")).read()
    declared = set(re.findall(r"\b(zpx_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    bound = {name for name, _, _ in zlib.SYMBOLS}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert hasattr(zlib.lib, name), name
    assert zlib.lib.zpx_abi_version() == 1


def test_error_names_are_the_zig_error_names(zlib):
    """codes 1..42 == the reference's error identifiers == the oracle's names"""
    for code in range(1, 43):
        assert zlib.lib.zpx_error_name(code).decode() == O.lib().zo_error_name(code).decode()
    ref = "/root/reference/src/jpeg/decoder.zig"
    if os.path.exists(ref):  # build container only
        zig_errors = set(re.findall(r"error\.([A-Za-z0-9]+)", open(ref).read()))
        names = {zlib.lib.zpx_error_name(c).decode() for c in range(1, 43)}
        # every error the decoder can return is representable (ShortHuffmanData is never raised)
        assert zig_errors - names <= {"ShortHuffmanData", "SOSMarkerNotFound", "InvalidStride", "InvalidPixData", "EndOfStream"}


def test_no_cpu_fallback_without_gpu(zlib):
    """Without a CUDA device the context cannot be created: there is no CPU decode path."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert zlib.lib.zpx_ctx_create(None, 0, C.byref(h)) == 101  # ZPX_E_NO_DEVICE
    assert not h.value


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "zpix_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                # no import, include, link or call of anything under oracle/
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s*[<\"].*oracle|libzpix_oracle|\bzo_[a-z_]+\s*\(", src, re.M), f


def _report(zlib, data):
    a = np.frombuffer(data, np.uint8)
    inf, rep = zlib.ZpxImageInfo(), zlib.ZpxParseReport()
    assert zlib.lib.zpx_parse_report_of(a.ctypes.data if a.size else None, a.size, C.byref(inf), C.byref(rep)) == 0
    return inf, rep


def test_parser_matches_oracle_on_fixtures(zlib, fixtures_dir):
    for name in sorted(os.listdir(fixtures_dir)):
        data = _read(fixtures_dir, name)
        img = O.decode(data)
        inf, rep = _report(zlib, data)
        prog = "progressive" in name or "separate.dc" in name
        assert rep.status == 0, name
        assert (inf.width, inf.height) == (img.width, img.height)
        assert inf.variant == img.variant, name
        assert bool(inf.progressive) == prog, name
        if img.variant == O.YCBCR:
            assert O.RATIO_NAMES[inf.subsample_ratio] == img.subsample_ratio
            assert (inf.y_stride, inf.c_stride) == (img.y_stride, img.c_stride)
            assert inf.native_len == img.pixels.size
        assert rep.pending_err == 0 and rep.trailing_err == 0


def test_parser_restart_intervals(zlib, fixtures_dir, golden_dir):
    inf, rep = _report(zlib, _read(fixtures_dir, "video-001.restart2.jpeg"))
    assert inf.restart_interval == 20 and rep.n_intervals == 4  # 70 MCUs / 20
    inf, rep = _report(zlib, _read(golden_dir, "padded_rst_issue28717.jpg"))
    assert rep.n_intervals > 1 and rep.pending_err == 0


def test_parser_error_kinds_match_oracle(zlib, fixtures_dir, golden_dir):
    """Header-level errors must be the oracle's error; errors the reference only finds while (or after)
    decoding entropy data must show up as pending/trailing errors of the same kind."""
    base = _read(fixtures_dir, "video-001.restart2.jpeg")
    cases = [b"", b"\xff", b"\xff\xd8", b"\x89PNG", _read(golden_dir, "fuzz_issue10413.bin")]
    cases += [base[:2816] + x + base[2816:] for x in (b"\xff\x03", b"\xff\xd5", b"\xff\xff\xd5", b"\x61")]
    g = _read(fixtures_dir, "video-005.gray.q50.jpeg")
    i = g.index(b"\xff\xda") + 2
    cases += [g[:k] for k in range(20, len(g), 97)] + [g[:k] for k in range(i, i + 10)]
    # header corruptions
    v = bytearray(_read(fixtures_dir, "video-001.jpeg"))
    sof = v.index(b"\xff\xc0")
    for off, val in [(4, 12), (9, 7), (11, 0x33), (11, 0x15), (12, 9)]:
        w = bytearray(v)
        w[sof + off] = val
        cases.append(bytes(w))
    dht = v.index(b"\xff\xc4")
    w = bytearray(v); w[dht + 4] = 0x25; cases.append(bytes(w))
    w = bytearray(v); w[dht + 3] = 0x00; w[dht + 2] = 0x05; cases.append(bytes(w))
    dqt = v.index(b"\xff\xdb")
    w = bytearray(v); w[dqt + 4] = 0x27; cases.append(bytes(w))
    w = bytearray(v); w[dqt + 4] = 0x20; cases.append(bytes(w))
    sos = v.index(b"\xff\xda")
    w = bytearray(v); w[sos + 5] = 99; cases.append(bytes(w))
    w = bytearray(v); w[sos + 6] = 0x40; cases.append(bytes(w))
    w = bytearray(v); w[sos + 3] = 0x09; cases.append(bytes(w))
    w = bytearray(v); w[2:4] = b"\xff\x01"; cases.append(bytes(w))
    w = bytearray(v); w[2:4] = b"\xff\xc9"; cases.append(bytes(w))
    for data in cases:
        try:
            O.decode(data)
            want = "ok"
        except O.OracleError as e:
            want = e.name
        inf, rep = _report(zlib, data)
        name = lambda c: zlib.lib.zpx_error_name(c).decode()
        if rep.status:
            assert name(rep.status) == want, (want, name(rep.status))
        elif rep.pending_err or rep.trailing_err:
            # found on the host, reported unless the device finds an earlier entropy error
            assert want != "ok"
            assert want in (name(rep.pending_err or rep.trailing_err), "MissingFF00", "BadHuffmanCode",
                            "ExcessiveDCComponent", "UninitializedHuffmanTable")
        else:
            # only the entropy decode (GPU) can fail these
            assert want in ("ok", "MissingFF00", "BadHuffmanCode", "ExcessiveDCComponent", "UninitializedHuffmanTable",
                            "UnsupportedColorModel"), want


ENTROPY_ERRORS = ("MissingFF00", "BadHuffmanCode", "ExcessiveDCComponent", "UninitializedHuffmanTable",
                  "UnexpectedHuffmanCode", "TooManyCoefficients")


def test_parser_header_fuzz_matches_oracle(zlib, fixtures_dir):
    """Seeded byte damage in the marker segments (SOF / DHT / DQT / DRI / SOS headers / APPn) of several fixtures:
    whenever the host parser rejects a file its error is the oracle's; when it accepts one, the oracle either
    decodes it or fails later, inside entropy data (the device's job) or with the error the parser queued."""
    rng = np.random.default_rng(20241018)
    name = lambda c: zlib.lib.zpx_error_name(c).decode()
    n_cases = n_rejected = 0
    for fname in ["video-001.jpeg", "video-001.q50.420.progressive.jpeg", "video-001.cmyk.jpeg", "video-001.restart2.jpeg",
                  "video-005.gray.q50.2x2.jpeg", "video-001.separate.dc.progression.jpeg", "video-001.rgb.jpeg"]:
        for d in header_damage(_read(fixtures_dir, fname), rng, 120):
            try:
                O.decode(d)
                want = "ok"
            except O.OracleError as e:
                want = e.name
            inf, rep = _report(zlib, d)
            n_cases += 1
            if rep.status:
                n_rejected += 1
                assert name(rep.status) == want, (fname, want, name(rep.status))
            elif rep.pending_err or rep.trailing_err:
                assert want != "ok", fname
                assert want in (name(rep.pending_err or rep.trailing_err),) + ENTROPY_ERRORS, (fname, want)
            else:
                assert want in ("ok", "UnsupportedColorModel") + ENTROPY_ERRORS or O.last_eob_carry(), (fname, want)
    assert n_cases == 840 and n_rejected > 100


def test_probe_is_decode_config(zlib, fixtures_dir):
    from zpix_b200 import jpeg

    for name in ("video-001.jpeg", "video-005.gray.jpeg", "video-001.cmyk.jpeg", "iceberg.jpg",
                 "video-001.progressive.jpeg"):
        data = _read(fixtures_dir, name)
        cfg = jpeg.decodeConfig(data)
        assert (cfg.width, cfg.height, cfg.color_model) == O.decode_config(data)
    with pytest.raises(jpeg.JpegError) as e:
        jpeg.decodeConfig(b"\x00\x01")
    assert e.value.name == "InvalidSOIMarker"
    assert jpeg.probeBuffer(_read(fixtures_dir, "iceberg.jpg")) and not jpeg.probeBuffer(b"\x89PNG")


def test_probe_reports_the_sizes_of_the_full_parse(zlib, fixtures_dir):
    """zpx_probe stops at the frame header (decodeConfig) but callers size their output buffers from it: variant,
    MCU grid, strides and byte lengths must be those of the full parse (decoder.zig:1258-1263, 1708-1783)."""
    L = zlib.lib
    for name in sorted(os.listdir(fixtures_dir)):
        data = np.frombuffer(_read(fixtures_dir, name), np.uint8)
        a, b = zlib.ZpxImageInfo(), zlib.ZpxImageInfo()
        rep = zlib.ZpxParseReport()
        assert L.zpx_probe(data.ctypes.data, data.size, C.byref(a)) == 0, name
        assert L.zpx_parse_report_of(data.ctypes.data, data.size, C.byref(b), C.byref(rep)) == 0
        for f in ("width", "height", "num_components", "variant", "subsample_ratio", "mxx", "myy", "y_stride", "c_stride",
                  "rgba_len", "native_len", "native_cb_off", "native_cr_off"):
            assert getattr(a, f) == getattr(b, f), (name, f)
        ref = O.decode(data.tobytes())
        assert a.native_len == ref.pixels.size, name


def test_unstuffing_plan_covers_every_interval(zlib, fixtures_dir):
    """host side of k0_unstuff: the pieces the parser cuts never split an FF 00 pair and add up to the unstuffed
    length (raw length minus stuffed zeros), for files with and without restart markers"""
    L = zlib.lib
    from tools import synth_jpeg as S
    datas = [_read(fixtures_dir, n) for n in ("video-001.restart2.jpeg", "video-001.jpeg", "iceberg.jpg")]
    datas.append(S.encode(70001, 640, 480, subsampling="4:2:0", restart_rows=1, quality=100))
    for d in datas:
        a = np.frombuffer(d, np.uint8)
        info, rep = zlib.ZpxImageInfo(), zlib.ZpxParseReport()
        assert L.zpx_parse_report_of(a.ctypes.data, a.size, C.byref(info), C.byref(rep)) == 0
        sos = d.index(b"\xff\xda")
        stuffed = d.count(b"\xff\x00", sos)
        assert rep.stuffed_bytes == stuffed
        assert rep.unstuffed_bytes + rep.stuffed_bytes + 2 * (rep.n_intervals - 1) <= rep.entropy_bytes + 2 * rep.n_intervals
        assert rep.n_pieces >= rep.n_intervals and rep.max_piece <= 16384 + 2
        assert rep.pieces_ok == 1


def test_unstuffing_kernel_index_logic_model():
    """device side of k0_unstuff, without a GPU: tools/k0_model.py restates the kernel's index logic lane by lane
    (linear output buffer, leftover move, shared head / tail vectors, boundary steps, zero padding of an interval's
    last piece) and compares it with a plain FF 00 -> FF replacement on random pieces of random alignment"""
    from tools import k0_model

    for seed in range(300):
        k0_model.test(seed)
    for seed in range(4):  # intervals of 6 - 40 KB cut at the product's piece size
        k0_model.test(seed, big=True)


def test_progressive_scan_script_classification(zlib, fixtures_dir):
    """Which progressive frames may take the lane-per-scan kernels (they apply correction bits as blind adds and keep
    the non-zero history in bit maps): every reference fixture (libjpeg's scripts) does; a frame whose script repeats a
    first pass over a band, or refines without lowering Al, does not -- and its unstuffing plan is still complete."""
    L = zlib.lib

    def report(d):
        a = np.frombuffer(d, np.uint8)
        info, rep = zlib.ZpxImageInfo(), zlib.ZpxParseReport()
        assert L.zpx_parse_report_of(a.ctypes.data, a.size, C.byref(info), C.byref(rep)) == 0
        return rep

    names = [n for n in sorted(os.listdir(fixtures_dir)) if "progressive" in n]
    assert len(names) >= 8
    for n in names:
        rep = report(_read(fixtures_dir, n))
        assert rep.status == 0 and rep.lane_script == 1 and rep.pieces_ok == 1, n
    assert report(_read(fixtures_dir, "video-001.jpeg")).lane_script == 0  # not progressive
    data = _read(fixtures_dir, "video-001.q50.420.progressive.jpeg")
    # walk the SOS headers: (offset of the Ah/Al byte, Ss, Ah, Al)
    pos, sos = 2, []
    while data[pos + 1] != 0xD9:
        ln = int.from_bytes(data[pos + 2:pos + 4], "big")
        if data[pos + 1] == 0xDA:
            q = pos + 2 + ln - 1
            sos.append((q, data[q - 2], data[q] >> 4, data[q] & 15))
            pos = q + 1
            while not (data[pos] == 0xFF and data[pos + 1] not in (0, *range(0xD0, 0xD8))):
                pos += 1
        else:
            pos += 2 + ln
    refinements = [s for s in sos if s[1] > 0 and s[2] > 0]
    assert refinements
    q, ss, ah, al = refinements[0]
    d = bytearray(data)
    d[q] = 0 << 4 | al  # the refinement becomes a second first pass over its band
    assert report(bytes(d)).lane_script == 0
    d[q] = (al + 2) << 4 | (al + 1)  # a refinement that repeats the Al of the first pass
    assert report(bytes(d)).lane_script == 0
    dc = [s for s in sos if s[1] == 0][0]
    d = bytearray(data)
    d[dc[0]] = 14  # Al = 14: a DC value shifted out of int16 range is the old kernel's business
    assert report(bytes(d)).lane_script == 0


def test_partition_rule(zlib):
    rng = np.random.default_rng(0)
    for n, nd in [(0, 1), (1, 8), (7, 2), (1024, 1), (1024, 2), (1024, 4), (1024, 8), (513, 8)]:
        w = rng.integers(1, 1000, n).astype(np.uint64)
        if n > 10:
            w[3] = 0  # an image that failed its header parse is not scheduled
        out = np.full(n, -7, np.int32)
        assert zlib.lib.zpx_partition(w.ctypes.data_as(C.POINTER(C.c_uint64)), n, nd,
                                      out.ctypes.data_as(C.POINTER(C.c_int32))) == 0
        sched = out[w > 0]
        assert (out[w == 0] == -1).all()
        if n:
            assert (np.diff(sched) >= 0).all() and sched.min() >= 0 and sched.max() < nd  # contiguous ranges
        if n >= 512:
            loads = np.array([w[(out == d)].sum() for d in range(nd)], float)
            assert loads.max() / loads.mean() < 1.05  # balanced by weight


def test_python_mirror_types():
    from zpix_b200 import color, image

    r = image.Rectangle.init(10, 20, 0, 0)
    assert (r.min.x, r.min.y, r.max.x, r.max.y) == (0, 0, 10, 20) and r.dX() == 10 and r.dY() == 20
    assert image.Point(3, 4).In(r) and not image.Point(10, 4).In(r)
    assert r.Intersect(image.Rectangle.init(5, 5, 50, 50)).dX() == 5
    assert r.Intersect(image.Rectangle.init(50, 50, 60, 60)) is None
    rng = np.random.default_rng(1)
    for y, cb, cr, k in rng.integers(0, 256, (500, 4)):
        y, cb, cr, k = int(y), int(cb), int(cr), int(k)
        assert tuple(v >> 8 for v in color.Color.fromYCbCr(y, cb, cr).toRGBA()) == O.ycbcr_to_rgba8(y, cb, cr)
        assert tuple(v >> 8 for v in color.Color.fromCMYK(y, cb, cr, k).toRGBA()) == O.cmyk_to_rgba8(y, cb, cr, k)
    assert color.Color.fromGray(7).toRGBA() == (0x0707, 0x0707, 0x0707, 0xFFFF)
    # at()/rgbaPixels() of the mirror types follow the reference's per-pixel path
    px = np.arange(4 * 3 * 2, dtype=np.uint8)
    img = image.Image("RGBA", image.RGBAImage(px, 12, image.Rectangle.init(0, 0, 3, 2)))
    assert np.array_equal(img.rgbaPixels(), px)
    yimg = image.YCbCrImage(np.full(64, 200, np.uint8), np.full(16, 90, np.uint8), np.full(16, 160, np.uint8), 8, 4,
                            image.YCbCrSubsample.Ratio420, image.Rectangle.init(0, 0, 8, 8), None)
    assert yimg.cOffset(5, 5) == 2 * 4 + 2
    assert tuple(v >> 8 for v in yimg.at(5, 5).toRGBA()) == O.ycbcr_to_rgba8(200, 90, 160)


def test_cpp_host_layer_compiles(tmp_path):
    """cpp/zpix.hpp (C++ mirror of the reference's API) builds against the header and the library."""
    exe = tmp_path / "zpix_demo"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "cpp", "zpix_demo.cpp"), "-o", str(exe),
                           "-L" + os.path.join(ROOT, "zpix_b200"), "-lzpixcuda",
                           "-Wl,-rpath," + os.path.join(ROOT, "zpix_b200")])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 64  # usage


def test_header_is_c99_clean(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "zpix_cuda.h"\nint main(void){ zpx_image_info i; (void)i; return zpx_abi_version() != 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c",
                           str(src), "-o", str(tmp_path / "t.o")])


WORKER = r'''
import os, sys, json
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
import bench
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
# weak scaling: every rank owns `per_rank` images; the job's value aggregates all ranks over the max time
per_rank = 5
pixels = per_rank * bench.W * bench.H
ms = bench.reduce_max(10.0 + 7.0 * rank, dist, "cpu")
value = bench.aggregate_value(pixels, world, ms)
seeds = bench.rank_seed_range(per_rank, rank)
out = [None] * world
dist.all_gather_object(out, (rank, ms, value, list(seeds), bench.host_thread_budget(world)))
# last phase of bench.py: a host-side (gloo) barrier, then every rank but 0 leaves and rank 0 goes on alone
# (the library's own multi-GPU scheduler is measured by one process over all devices)
g = dist.new_group(backend="gloo")
dist.barrier(group=g)
dist.destroy_process_group()
if rank == 0:
    import time
    time.sleep(0.5)
    print(json.dumps(out))
'''


def test_bench_multi_rank_plumbing_gloo(tmp_path):
    """world_size 2 over gloo: max-over-ranks timing, whole-job aggregate, per-rank shards."""
    script = tmp_path / "w.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29571")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29571", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    import json

    line = [l for l in r.stdout.splitlines() if l.startswith("[[")][-1]
    res = json.loads(line)
    assert [x[0] for x in res] == [0, 1]
    assert all(abs(x[1] - 17.0) < 1e-6 for x in res)  # max over ranks
    want = 2 * 5 * 1920 * 1080 / 1e6 / 0.017
    assert all(abs(x[2] - want) / want < 1e-9 for x in res)
    assert res[0][3] == res[1][3]  # every rank decodes its own copy of the same cfg2 batch
    assert res[0][4] == max(2, (os.cpu_count() or 1) // 2)  # the ranks of a box share its host cores
