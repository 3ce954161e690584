# Top-level build: the product library, the oracle, the test harnesses.
#
#   make            -> bjxa_b200/lib/libbjxa_b200.so  (CUDA sm_100a + host C)
#   make oracle     -> oracle/_build, oracle/_ref     (test infrastructure)
#   make emul       -> tests/_build/libxa_emul.so     (CPU single-stepper of
#                                                      the tile code; tests only)
#   make dropin     -> oracle/_ref/bjxa_dropin ...    (the reference's own CLI
#                      and API test, compiled from /root/reference and linked
#                      against OUR library: the drop-in acceptance binaries)

NVCC     ?= nvcc
CC       ?= gcc
CXX      ?= g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v $(XA_DEFS)
CFLAGS   := -std=c99 -O2 -fPIC -Wall -Wextra -Wno-unused-parameter
REFERENCE ?= /root/reference

SRC      := bjxa_b200/csrc
LIBDIR   := bjxa_b200/lib
OBJDIR   := build
LIB      := $(LIBDIR)/libbjxa_b200.so
HDRS     := $(SRC)/xa_core.h $(SRC)/xa_tile.h $(SRC)/xa_plan.h $(SRC)/xa_walk.h $(SRC)/bjxa_internal.h include/bjxa.h include/bjxa_batch.h

.PHONY: all lib oracle emul dropin clean
all: lib

lib: $(LIB)

$(OBJDIR)/xa_kernels.o: $(SRC)/xa_kernels.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c -o $@ $< 2> $(OBJDIR)/ptxas.log || (cat $(OBJDIR)/ptxas.log; false)

$(OBJDIR)/bjxa_host.o: $(SRC)/bjxa_host.c $(SRC)/bjxa_internal.h include/bjxa.h include/bjxa_batch.h
	@mkdir -p $(OBJDIR)
	$(CC) $(CFLAGS) -c -o $@ $<

$(OBJDIR)/bjxa_corpus.o: $(SRC)/bjxa_corpus.c $(SRC)/bjxa_internal.h include/bjxa.h include/bjxa_batch.h
	@mkdir -p $(OBJDIR)
	$(CC) $(CFLAGS) -c -o $@ $<

OBJS     := $(OBJDIR)/xa_kernels.o $(OBJDIR)/bjxa_host.o $(OBJDIR)/bjxa_corpus.o
LINK      = $(NVCC) $(ARCH) -shared -o $@ $(OBJS) -Xlinker --version-script=$(SRC)/libbjxa.map \
	    -cudart static -lpthread -ldl -lrt

# The same objects under two names: libbjxa_b200.so for callers that ask for this
# backend by name, and libbjxa.so.0 -- the soname the reference installs
# (/root/reference/Makefile.am:28, -version-info 2:0:2 -> libbjxa.so.0.2.0) -- so
# that a program ALREADY LINKED against the reference picks the backend up through
# the library path, no relink; plus libbjxa.so for -lbjxa and a bjxa.pc like the
# reference's (/root/reference/bjxa.pc.in, Makefile.am:43).  Plain copies rather
# than symlinks: the directory travels to machines by tools that may not keep links.
$(LIB): $(OBJS) $(SRC)/libbjxa.map bjxa_b200/bjxa.pc.in
	@mkdir -p $(LIBDIR)/pkgconfig
	$(LINK) -Xlinker -soname=libbjxa_b200.so
	$(NVCC) $(ARCH) -shared -o $(LIBDIR)/libbjxa.so.0 $(OBJS) -Xlinker --version-script=$(SRC)/libbjxa.map \
	    -Xlinker -soname=libbjxa.so.0 -cudart static -lpthread -ldl -lrt
	cp -f $(LIBDIR)/libbjxa.so.0 $(LIBDIR)/libbjxa.so
	sed -e 's|@libdir@|$(abspath $(LIBDIR))|' -e 's|@includedir@|$(abspath include)|' \
	    bjxa_b200/bjxa.pc.in > $(LIBDIR)/pkgconfig/bjxa.pc

oracle:
	$(MAKE) -s -C oracle all

emul: tests/_build/libxa_emul.so

tests/_build/libxa_emul.so: tests/emul/xa_emul.cc $(HDRS)
	@mkdir -p tests/_build
	$(CXX) -std=c++17 -O1 -g -fPIC -Wall -Wno-unknown-pragmas -shared -o $@ tests/emul/xa_emul.cc

# The reference's CLI and API test, unmodified, against the product library.
ifneq ($(wildcard $(REFERENCE)/src/bjxa.c),)
REF_CLI  := $(REFERENCE)/src/bjxa.c $(REFERENCE)/src/bjxa_decode.c $(REFERENCE)/src/bjxa_encode.c
REF_CF   := -std=c99 -D_POSIX_C_SOURCE=200809L -D_XOPEN_SOURCE=600 -O2 -Ioracle/_ref -I$(REFERENCE)/src
DROPLINK := -L$(LIBDIR) -lbjxa_b200 -Wl,-rpath,'$$ORIGIN/../../$(LIBDIR)'
dropin: lib oracle
	$(CC) $(REF_CF) -o oracle/_ref/bjxa_dropin $(REF_CLI) $(DROPLINK)
	$(CC) $(REF_CF) -DBJXA_SINGLE_PASS -o oracle/_ref/bjxa_dropin_single_pass $(REF_CLI) $(DROPLINK)
	$(CC) $(REF_CF) -o oracle/_ref/test_api_dropin $(REFERENCE)/test/test_libbjxa_api.c $(DROPLINK)
	# the no-relink case: the reference's library under its own soname, and its CLI
	# linked against THAT (no rpath); tests swap the library path to ours
	@mkdir -p oracle/_ref/refso
	$(CC) $(REF_CF) -fPIC -shared -Wl,-soname,libbjxa.so.0 -Wl,--version-script=$(REFERENCE)/src/libbjxa.map \
	    -o oracle/_ref/refso/libbjxa.so.0 $(REFERENCE)/src/libbjxa.c
	cp -f oracle/_ref/refso/libbjxa.so.0 oracle/_ref/refso/libbjxa.so
	$(CC) $(REF_CF) -o oracle/_ref/bjxa_ref_dyn $(REF_CLI) -Loracle/_ref/refso -lbjxa
else
dropin:
	@echo "dropin: $(REFERENCE) not present; using prebuilt oracle/_ref binaries if any"
endif

# The reference's habit (configure.ac:41-43,67-75: --enable-asan / --enable-ubsan):
# the host C of the library, the oracle and the CPU single-stepper of the tile code
# under AddressSanitizer + UndefinedBehaviorSanitizer, the whole CPU test suite run
# against them.  (The kernels cannot be run under compute-sanitizer on this pool:
# the tool is closed there; tests/emul single-steps their tile code on the CPU,
# which is what the sanitizers then see.)
SAN      := -fsanitize=address,undefined -fno-omit-frame-pointer -g
SANDIR   := $(OBJDIR)/san
# a compiler that ships the sanitizer runtimes (the distribution's)
SANCC    ?= /usr/bin/gcc
SANCXX   ?= /usr/bin/g++
.PHONY: sanitize
sanitize: $(OBJDIR)/xa_kernels.o
	@mkdir -p $(SANDIR)
	$(SANCC) $(CFLAGS) $(SAN) -c -o $(SANDIR)/bjxa_host.o $(SRC)/bjxa_host.c
	$(SANCC) $(CFLAGS) $(SAN) -c -o $(SANDIR)/bjxa_corpus.o $(SRC)/bjxa_corpus.c
	$(NVCC) $(ARCH) -shared -o $(SANDIR)/libbjxa_b200.so $(OBJDIR)/xa_kernels.o $(SANDIR)/bjxa_host.o \
	    $(SANDIR)/bjxa_corpus.o -Xlinker --version-script=$(SRC)/libbjxa.map -cudart static \
	    -lpthread -ldl -lrt
	$(SANCXX) -std=c++17 -O1 -g $(SAN) -fPIC -Wall -Wno-unknown-pragmas -shared -o $(SANDIR)/libxa_emul.so tests/emul/xa_emul.cc
	$(SANCC) -std=c99 -O1 -g $(SAN) -fPIC -shared -o $(SANDIR)/libbjxa_oracle.so oracle/bjxa_oracle.c
	LD_PRELOAD="$$($(SANCC) -print-file-name=libasan.so) $$($(SANCC) -print-file-name=libubsan.so)" \
	    ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 UBSAN_OPTIONS=halt_on_error=1:print_stacktrace=1 \
	    BJXA_B200_LIB=$(abspath $(SANDIR))/libbjxa_b200.so XA_EMUL_SO=$(abspath $(SANDIR))/libxa_emul.so \
	    BJXA_ORACLE_SO=$(abspath $(SANDIR))/libbjxa_oracle.so \
	    python -m pytest tests -q -m "not gpu" -p no:cacheprovider 2>&1 | tee profiles/sanitize_host.log | tail -5

clean:
	rm -rf $(OBJDIR) $(LIBDIR) tests/_build
	$(MAKE) -s -C oracle clean
