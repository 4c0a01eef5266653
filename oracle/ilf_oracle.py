"""ctypes door onto oracle/liboracle.so (the CPU restatement, oracle/ilf_oracle.c).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
All functions take/return numpy arrays; pictures are dicts {"y","cb","cr"} of int16 2-D arrays.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
    return _LIB


def _p(a, t=C.c_void_p):
    return None if a is None else a.ctypes.data_as(t)


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def deblock(pic, bd_luma, bd_chroma, ctu_log2, params_bytes, info, info_chroma=None, mv16=None, mv32=None, ctu_slice=None):
    """LoopFilter::loopFilterPic on a copy of `pic`; returns the filtered picture."""
    y, cb, cr = (np.array(pic[k], dtype=np.int16, order="C") for k in ("y", "cb", "cr"))
    h, w = y.shape
    pb = _c(np.frombuffer(bytes(params_bytes), dtype=np.uint8), np.uint8)
    info = _c(info, np.uint32); info_chroma = _c(info_chroma, np.uint32)
    mv16 = _c(mv16, np.int16); mv32 = _c(mv32, np.int32); ctu_slice = _c(ctu_slice, np.uint8)
    f = lib().ilf_oracle_deblock
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_void_p, C.c_ssize_t] + [C.c_int] * 5 + [C.c_void_p] * 6
    rc = f(_p(y), y.shape[1], _p(cb), _p(cr), cb.shape[1], w, h, bd_luma, bd_chroma, ctu_log2, _p(pb), _p(info),
           _p(info_chroma), _p(mv16), _p(mv32), _p(ctu_slice))
    assert rc == 0
    return {"y": y, "cb": cb, "cr": cr}


def bs_map(width, height, params_bytes, info, mv16=None, mv32=None):
    pb = _c(np.frombuffer(bytes(params_bytes), dtype=np.uint8), np.uint8)
    info = _c(info, np.uint32); mv16 = _c(mv16, np.int16); mv32 = _c(mv32, np.int32)
    out = np.zeros((2, height // 4, width // 4), dtype=np.uint8)
    f = lib().ilf_oracle_bs_map
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4
    assert f(_p(out), width, height, _p(pb), _p(info), _p(mv16), _p(mv32)) == 0
    return out


def _planes3(pic):
    arrs = [np.array(pic[k], dtype=np.int16, order="C") for k in ("y", "cb", "cr")]
    ptrs = (C.c_void_p * 3)(*[a.ctypes.data for a in arrs])
    strides = (C.c_ssize_t * 3)(*[a.shape[1] for a in arrs])
    return arrs, ptrs, strides


def sao(pic, bd_luma, bd_chroma, ctu_log2, sao_ctus):
    """SAOProcess with resolved per-CTU parameters (uint8 array [num_ctus, 32] of ilf_sao_ctu)."""
    src, sp, ss = _planes3(pic)
    dst, dp, ds = _planes3(pic)
    ctus = _c(sao_ctus, np.uint8)
    h, w = src[0].shape
    f = lib().ilf_oracle_sao
    f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_void_p]
    assert f(sp, ss, dp, ds, w, h, bd_luma, bd_chroma, ctu_log2, _p(ctus)) == 0
    return {"y": dst[0], "cb": dst[1], "cr": dst[2]}


def alf(pic, bd_luma, bd_chroma, ctu_log2, alf_params_bytes, ctu_enable):
    src, sp, ss = _planes3(pic)
    dst, dp, ds = _planes3(pic)
    pb = _c(np.frombuffer(bytes(alf_params_bytes), dtype=np.uint8), np.uint8)
    en = _c(ctu_enable, np.uint8)
    h, w = src[0].shape
    f = lib().ilf_oracle_alf
    f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_void_p] * 2
    assert f(sp, ss, dp, ds, w, h, bd_luma, bd_chroma, ctu_log2, _p(pb), _p(en)) == 0
    return {"y": dst[0], "cb": dst[1], "cr": dst[2]}


def alf_classify(y, bd_luma):
    y = np.ascontiguousarray(y, dtype=np.int16)
    h, w = y.shape
    out = np.zeros((h // 4, w // 4), dtype=np.uint8)
    f = lib().ilf_oracle_alf_classify
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_int, C.c_void_p]
    assert f(_p(y), w, w, h, bd_luma, _p(out)) == 0
    return out


def sao_stats(rec, org, bd_luma, bd_chroma, ctu_log2, avail):
    """EncSampleAdaptiveOffset::getStatistics (no SaoCtuBoundary): int64 array [num_ctus, 3, 5, 64] (diff[32], count[32])."""
    r, rp, rs = _planes3(rec)
    o, op, os_ = _planes3(org)
    h, w = r[0].shape
    ctu = 1 << ctu_log2
    n = ((w + ctu - 1) // ctu) * ((h + ctu - 1) // ctu)
    av = _c(avail, np.uint8)
    assert av.size == n
    out = np.zeros((n, 3, 5, 64), dtype=np.int64)
    f = lib().ilf_oracle_sao_stats
    f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_void_p, C.c_void_p]
    assert f(rp, rs, op, os_, w, h, bd_luma, bd_chroma, ctu_log2, _p(av), _p(out)) == 0
    return out


def sao_stats_block(src, org, x0, y0, w, h, bd, is_chroma, avail6):
    """One block of 2-D arrays src/org at (x0, y0), explicit flags (bit0 L, 1 R, 2 A, 3 B, 4 AL, 5 AR): int64 [5, 64]."""
    src = np.ascontiguousarray(src, np.int16); org = np.ascontiguousarray(org, np.int16)
    out = np.zeros((5, 64), dtype=np.int64)
    f = lib().ilf_oracle_sao_stats_block
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_void_p]
    sp = src.ctypes.data + 2 * (y0 * src.shape[1] + x0)
    gp = org.ctypes.data + 2 * (y0 * org.shape[1] + x0)
    assert f(sp, src.shape[1], gp, org.shape[1], w, h, bd, int(is_chroma), avail6, _p(out)) == 0
    return out


def picture_hash(pic, bd_luma, bd_chroma, kind):
    """calcCRC / calcChecksum (PicYuvMD5.cpp:130-175): [Y, Cb, Cr] values; kind = "crc" or "checksum"."""
    f = lib().ilf_oracle_crc if kind == "crc" else lib().ilf_oracle_checksum
    f.restype = C.c_uint
    f.argtypes = [C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_int]
    out = []
    for k, bd in (("y", bd_luma), ("cb", bd_chroma), ("cr", bd_chroma)):
        a = np.ascontiguousarray(pic[k], dtype=np.int16)
        out.append(int(f(_p(a), a.shape[1], a.shape[1], a.shape[0], bd)))
    return out


def extend_border(plane, xmargin, ymargin):
    """Picture::extendPicBorder of one plane: returns the (h + 2 ymargin, w + 2 xmargin) array with replicated borders."""
    plane = np.ascontiguousarray(plane, dtype=np.int16)
    h, w = plane.shape
    big = np.full((h + 2 * ymargin, w + 2 * xmargin), -1, np.int16)
    big[ymargin:ymargin + h, xmargin:xmargin + w] = plane
    f = lib().ilf_oracle_extend_border
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_int, C.c_int]
    f(big.ctypes.data + 2 * (ymargin * big.shape[1] + xmargin), big.shape[1], w, h, xmargin, ymargin)
    return big


ALF_STATS_WORDS = 25 * 105 + 36 + 36
ALF_5X5_IN_7X7 = (2, 5, 6, 7, 10, 11, 12)   # coefficient k of the 5x5 shape sits at this index of the 7x7 shape


def alf_stats(rec, org, bd_luma, ctu_log2):
    """EncAdaptiveLoopFilter::deriveStatsForFiltering: int64 [num_ctus, ALF_STATS_WORDS] (luma [25][105] 7x7, Cb [36], Cr [36] 5x5;
    a record = E upper triangle row-major, y, pixAcc)."""
    r, rp, rs = _planes3(rec)
    o, op, os_ = _planes3(org)
    h, w = r[0].shape
    ctu = 1 << ctu_log2
    n = ((w + ctu - 1) // ctu) * ((h + ctu - 1) // ctu)
    out = np.zeros((n, ALF_STATS_WORDS), dtype=np.int64)
    f = lib().ilf_oracle_alf_stats
    f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_void_p]
    assert f(rp, rs, op, os_, w, h, bd_luma, ctu_log2, _p(out)) == 0
    return out


def alf_stats_unpack(rec_words, n):
    """(E [n, n] symmetric, y [n], pixAcc) from one record of n (n + 1) / 2 + n + 1 words."""
    E = np.zeros((n, n), np.int64)
    i = 0
    for k in range(n):
        for l in range(k, n):
            E[k, l] = E[l, k] = rec_words[i]
            i += 1
    return E, np.array(rec_words[i:i + n]), int(rec_words[i + n])
