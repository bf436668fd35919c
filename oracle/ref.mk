# ref.mk -- compiles the UNMODIFIED reference (SKIRT 9) from the sources where they lie under $(REF) into
# oracle/_ref/ (git-ignored, NOT gpurun-ignored: the binary travels to the GPU box).  No reference source is copied
# and the reference's own cmake build is not used: every translation unit is handed to gcc/g++ directly.
# The only files written besides objects are the two headers cmake would have generated for the log banner
# (version string and build time stamp, SMILE/build/CMakeLists.txt:31-61) and a placeholder resource file so that
# FilePaths::findResources (SKIRT/core/FilePaths.cpp:108-137) finds its <exe>/../../../git/SKIRT/resources directory
# on machines where /root/reference does not exist.
#   make -f ref.mk -j8            (about 12 CPU-minutes)
REF  ?= /root/reference
OUT  ?= _ref
CXX  ?= g++
CC   ?= gcc
CXXFLAGS = -O3 -DNDEBUG -std=c++14 -pthread -w
CFLAGS   = -O3 -DNDEBUG -DFF_NO_UNISTD_H=1 -w
INC = -I$(OUT)/gen $(addprefix -I$(REF)/,SMILE/fundamentals SMILE/schema SMILE/serialize SMILE/build \
      SKIRT/utils SKIRT/mpi SKIRT/core SKIRT/fitsio SKIRT/voro SKIRT/tetgen)

CPPDIRS = SMILE/fundamentals SMILE/schema SMILE/serialize SMILE/build SKIRT/utils SKIRT/mpi SKIRT/core SKIRT/main
CPPSRC  = $(foreach d,$(CPPDIRS),$(wildcard $(REF)/$(d)/*.cpp))
CCSRC   = $(wildcard $(REF)/SKIRT/voro/*.cc)
CXXSRC  = $(wildcard $(REF)/SKIRT/tetgen/*.cxx)
CSRC    = $(wildcard $(REF)/SKIRT/fitsio/*.c)
OBJ = $(patsubst $(REF)/%.cpp,$(OUT)/obj/%.o,$(CPPSRC)) $(patsubst $(REF)/%.cc,$(OUT)/obj/%.o,$(CCSRC)) \
      $(patsubst $(REF)/%.cxx,$(OUT)/obj/%.o,$(CXXSRC)) $(patsubst $(REF)/%.c,$(OUT)/obj/%.o,$(CSRC))
EXE = $(OUT)/release/SKIRT/main/skirt

all: $(EXE) $(OUT)/git/SKIRT/resources/PLACEHOLDER.txt

$(OUT)/gen/version.h:
	@mkdir -p $(OUT)/gen
	@printf '#ifndef VERSION_HPP\n#define VERSION_HPP\n#define PROJECT_VERSION "v9.0"\n#endif\n' > $@
	@printf '#ifndef TIMESTAMP_H\n#define TIMESTAMP_H\n#define BUILD_DATE "oracle"\n#define BUILD_TIME "ref.mk"\n#define COMMIT_HASH "1facef2"\n#endif\n' > $(OUT)/gen/timestamp.h

$(OUT)/git/SKIRT/resources/PLACEHOLDER.txt:
	@mkdir -p $(dir $@)
	@echo "placeholder so that the reference finds a built-in resource directory; no resource packs are installed" > $@

$(OUT)/obj/%.o: $(REF)/%.cpp $(OUT)/gen/version.h
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) $(INC) -c $< -o $@
$(OUT)/obj/%.o: $(REF)/%.cc
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) $(INC) -c $< -o $@
$(OUT)/obj/%.o: $(REF)/%.cxx
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) -DTETLIBRARY $(INC) -c $< -o $@
$(OUT)/obj/%.o: $(REF)/%.c
	@mkdir -p $(dir $@)
	$(CC) $(CFLAGS) -c $< -o $@

$(EXE): $(OBJ)
	@mkdir -p $(dir $@)
	$(CXX) -pthread -o $@ $(OBJ)
	@echo built $@

.PHONY: all
