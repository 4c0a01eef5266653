# oracle/ref.mk -- TEST INFRASTRUCTURE, not product code.
#
# Compiles the UNMODIFIED reference (VTM 2.1, /root/reference) from the sources where they lie,
# with g++ directly (the reference's own CMake system is NOT run: it writes bin/ and lib/ into its
# source tree).  Every output goes to oracle/_ref/ (git-ignored, but it
# travels to the GPU box with the gpurun snapshot).  No reference source is copied into the repo.
#
# Products:
#   oracle/_ref/libvtm.a          CommonLib + DecoderLib + EncoderLib + Utilities + libmd5
#   oracle/_ref/DecoderApp        stock decoder  (hash-SEI parity check, CPU baseline)
#   oracle/_ref/EncoderApp        stock encoder  (makes the committed bitstreams in tests/golden/)
#   (oracle/Makefile builds the capture/shim decoders on top of libvtm.a)
#
# Flags mirror the reference's CMake configuration with -DENABLE_VTM=ON
# (CMakeLists.txt:99-101 -msse4.1; source/Lib/CommonLib/CMakeLists.txt:46-48 BMS_TOOLS=0,
#  :72-75 ENABLE_*_PARALLELISM=0, :84-98 per-directory SIMD flags).  "-include cstdint -include
# limits" is needed by gcc 13 (TypeDef.h:375 uses uint32_t without the header).

REF      ?= /root/reference
OUT      ?= $(CURDIR)/_ref
SRC      := $(REF)/source
CXX      ?= g++
DEFS     := -DBMS_TOOLS=0 -DENABLE_SPLIT_PARALLELISM=0 -DENABLE_WPP_PARALLELISM=0
INCS     := -I$(SRC)/Lib -I$(SRC)/Lib/CommonLib -I$(SRC)/Lib/CommonLib/x86 -I$(SRC)/Lib/libmd5 \
            -I$(SRC)/Lib/DecoderLib -I$(SRC)/Lib/EncoderLib -I$(SRC)/Lib/Utilities -I$(OUT)/gen
CXXFLAGS := -std=c++11 -O3 -fPIC -msse4.1 -w -include cstdint -include limits $(DEFS) $(INCS)

LIBDIRS  := Lib/CommonLib Lib/CommonLib/x86 Lib/CommonLib/x86/sse41 Lib/CommonLib/x86/avx \
            Lib/CommonLib/x86/avx2 Lib/libmd5 Lib/DecoderLib Lib/EncoderLib Lib/Utilities
LIBSRCS  := $(foreach d,$(LIBDIRS),$(wildcard $(SRC)/$(d)/*.cpp))
LIBOBJS  := $(patsubst $(SRC)/%.cpp,$(OUT)/obj/%.o,$(LIBSRCS))
DECSRCS  := $(wildcard $(SRC)/App/DecoderApp/*.cpp)
DECOBJS  := $(patsubst $(SRC)/%.cpp,$(OUT)/obj/%.o,$(DECSRCS))
ENCSRCS  := $(wildcard $(SRC)/App/EncoderApp/*.cpp)
ENCOBJS  := $(patsubst $(SRC)/%.cpp,$(OUT)/obj/%.o,$(ENCSRCS))

.PHONY: all lib clean
all: $(OUT)/DecoderApp $(OUT)/EncoderApp
lib: $(OUT)/libvtm.a

$(OUT)/obj/Lib/CommonLib/x86/sse41/%.o: EXTRA := -msse4.1 -DUSE_SSE41
$(OUT)/obj/Lib/CommonLib/x86/avx/%.o:   EXTRA := -mavx -DUSE_AVX
$(OUT)/obj/Lib/CommonLib/x86/avx2/%.o:  EXTRA := -mavx2 -DUSE_AVX2

# svnrevision.h is the one generated header (cmake/modules/GetSVN.cmake); outside a subversion
# checkout that script writes an EMPTY file, which is what we create here.
$(OUT)/gen/svnrevision.h:
	@mkdir -p $(dir $@)
	@: > $@

$(OUT)/obj/%.o: $(SRC)/%.cpp $(OUT)/gen/svnrevision.h
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) $(EXTRA) -c $< -o $@

$(OUT)/libvtm.a: $(LIBOBJS)
	@rm -f $@
	ar rcs $@ $(LIBOBJS)

$(OUT)/DecoderApp: $(DECOBJS) $(OUT)/libvtm.a
	$(CXX) -o $@ $(DECOBJS) $(OUT)/libvtm.a -lpthread

$(OUT)/EncoderApp: $(ENCOBJS) $(OUT)/libvtm.a
	$(CXX) -o $@ $(ENCOBJS) $(OUT)/libvtm.a -lpthread

clean:
	rm -rf $(OUT)/obj $(OUT)/libvtm.a $(OUT)/DecoderApp $(OUT)/EncoderApp
