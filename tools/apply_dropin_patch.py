#!/usr/bin/env python3
"""Applies the drop-in patch of INTEGRATION.md to a COPY of the reference tree (VTM 2.1): the ILF_B200 CMake option and the
`#if !ILF_B200` guards around the five entry points the host shim supplies.  Edits in place -- never run it on /root/reference.

  tools/apply_dropin_patch.py <copy of the reference> <this repository>

Then:  cmake <copy> -DENABLE_VTM=ON -DILF_B200=ON -DILF_B200_ROOT=<this repository> && make DecoderApp EncoderApp
(tools/cmake_dropin_build.sh does all of it in a scratch directory and checks that the binaries reach the library.)"""
import os
import re
import sys


def guard_function(path, signature_regex):
    """Wrap the definition that starts with `signature_regex` (at column 0) and ends at its matching closing brace."""
    src = open(path).read()
    m = re.search(signature_regex, src, flags=re.M)
    if not m:
        raise SystemExit(f"{path}: cannot find {signature_regex}")
    start = m.start()
    # an #if K0238... / #else / #endif block may wrap two alternative signatures: start the guard before it
    before = src.rfind("\n#if", 0, start)
    if before != -1 and src[before:start].count("\n") <= 2 and "#endif" not in src[before:start]:
        start = before + 1
        m2 = list(re.finditer(signature_regex, src, flags=re.M))
        body_from = m2[-1].end()
    else:
        body_from = m.end()
    i = src.index("{", body_from)
    depth = 0
    while True:
        c = src[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                break
        i += 1
    end = i + 1
    src = src[:start] + "#if !ILF_B200   // libilf_b200: the host shim supplies this entry point\n" + src[start:end] + "\n#endif\n" + src[end:]
    open(path, "w").write(src)


def main():
    ref, repo = os.path.abspath(sys.argv[1]), os.path.abspath(sys.argv[2])
    if ref.startswith("/root/reference"):
        raise SystemExit("refusing to edit /root/reference: work on a copy")
    lib = os.path.join(ref, "source", "Lib")
    guard_function(os.path.join(lib, "CommonLib", "LoopFilter.cpp"), r"^void LoopFilter::loopFilterPic\s*\(")
    guard_function(os.path.join(lib, "CommonLib", "SampleAdaptiveOffset.cpp"), r"^void SampleAdaptiveOffset::SAOProcess\s*\(")
    guard_function(os.path.join(lib, "CommonLib", "AdaptiveLoopFilter.cpp"), r"^void AdaptiveLoopFilter::ALFProcess\s*\(")
    guard_function(os.path.join(lib, "EncoderLib", "EncSampleAdaptiveOffset.cpp"), r"^void EncSampleAdaptiveOffset::SAOProcess\s*\(")
    guard_function(os.path.join(lib, "EncoderLib", "EncAdaptiveLoopFilter.cpp"), r"^void EncAdaptiveLoopFilter::ALFProcess\s*\(")
    # top-level option
    p = os.path.join(ref, "CMakeLists.txt")
    s = open(p).read()
    opt = '''
# libilf_b200: deblocking / SAO / ALF on an NVIDIA B200 (INTEGRATION.md of the library)
option( ILF_B200 "run deblocking/SAO/ALF on an NVIDIA B200 through libilf_b200" OFF )
set( ILF_B200_ROOT "" CACHE PATH "checkout of the libilf_b200 repository" )
if( ILF_B200 )
  add_definitions( -DILF_B200=1 )
  include_directories( ${ILF_B200_ROOT}/include ${ILF_B200_ROOT}/vvcsoftware_vtm_b200/shim )
  link_directories( ${ILF_B200_ROOT}/vvcsoftware_vtm_b200 )
else()
  add_definitions( -DILF_B200=0 )
endif()
'''
    anchor = "# set c++11"
    assert anchor in s
    s = s.replace(anchor, opt + "\n" + anchor, 1)
    open(p, "w").write(s)
    # CommonLib: packer + shim; EncoderLib: encoder shim; both link the library
    for libname, files in (("CommonLib", ["ilf_pack.cpp", "ilf_shim.cpp"]), ("EncoderLib", ["ilf_shim_enc.cpp"])):
        p = os.path.join(lib, libname, "CMakeLists.txt")
        s = open(p).read()
        anchor = "add_library( ${LIB_NAME} STATIC"
        assert anchor in s
        extra = "if( ILF_B200 )\n" + "".join(f"  list( APPEND SRC_FILES ${{ILF_B200_ROOT}}/vvcsoftware_vtm_b200/shim/{f} )\n" for f in files) + "endif()\n"
        s = s.replace(anchor, extra + anchor, 1)
        s += "\nif( ILF_B200 )\n  target_link_libraries( ${LIB_NAME} ilf_b200 )\nendif()\n"
        open(p, "w").write(s)
    print("patched", ref)


if __name__ == "__main__":
    main()
