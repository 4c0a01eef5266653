"""Reader/writer for the .ilfcap container written by oracle/capture_hook.cpp.

Layout: 8-byte magic "ILFCAP1\\0", then records
  name[24] | dtype u8 | ndim u8 | pad[6] | dims u32[3] | nbytes u64 | data
dtype: 0 u8, 1 i16, 2 i32, 3 u32.  A capture holds one picture: geometry ("geom"), the flat side
information of include/ilf_b200.h ("db_*", "sao_*", "alf_*") and the planes before deblocking ("pre_*")
and after each reference stage ("dbk_*", "sao_*", "alf_*" with suffix _y/_cb/_cr).
"""
import struct
import numpy as np

_DT = {0: np.uint8, 1: np.int16, 2: np.int32, 3: np.uint32}
_DTI = {np.dtype(v): k for k, v in _DT.items()}
GEOM_FIELDS = ("width", "height", "bd_luma", "bd_chroma", "ctu_log2", "poc", "slice_type", "use_sao",
               "use_alf", "dual_tree", "deblock_us", "sao_us", "alf_us", "num_slices")


def load(path):
    out = {}
    with open(path, "rb") as fh:
        data = fh.read()
    assert data[:8] == b"ILFCAP1\0", "not an ilfcap file"
    pos = 8
    while pos < len(data):
        name = data[pos:pos + 24].split(b"\0")[0].decode()
        dtype, ndim = data[pos + 24], data[pos + 25]
        dims = struct.unpack_from("<3I", data, pos + 32)
        (nbytes,) = struct.unpack_from("<Q", data, pos + 44)
        pos += 52
        arr = np.frombuffer(data, dtype=_DT[dtype], count=nbytes // np.dtype(_DT[dtype]).itemsize, offset=pos)
        out[name] = arr.reshape(dims[:ndim]).copy()
        pos += nbytes
    if "geom" in out:
        out["geom"] = dict(zip(GEOM_FIELDS, (int(v) for v in out["geom"])))
    return out


def save(path, arrays):
    with open(path, "wb") as fh:
        fh.write(b"ILFCAP1\0")
        for name, arr in arrays.items():
            if name == "geom" and isinstance(arr, dict):
                arr = np.array([arr.get(k, 0) for k in GEOM_FIELDS] + [0, 0], dtype=np.int32)
            arr = np.ascontiguousarray(arr)
            dims = list(arr.shape)[:3] + [1] * (3 - arr.ndim)
            fh.write(name.encode().ljust(24, b"\0"))
            fh.write(bytes([_DTI[arr.dtype], arr.ndim, 0, 0, 0, 0, 0, 0]))
            fh.write(struct.pack("<3I", *dims))
            fh.write(struct.pack("<Q", arr.nbytes))
            fh.write(arr.tobytes())
