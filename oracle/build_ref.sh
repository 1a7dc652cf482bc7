#!/bin/bash
# Builds the UNMODIFIED reference (POV-Ray 3.8.0-alpha) from the sources where they lie under
# /root/reference into oracle/_ref/ (git-ignored; travels to the GPU box).  This is TEST / BASELINE
# infrastructure only: nothing under povray_b200/ may link or execute it.
#
# Not the reference's own build system (autotools is absent): a hand-written config.h plus direct
# g++ invocations, following SURVEY.md Appendix C.  Two variants:
#   parity : -O2 -fno-fast-math -ffp-contract=off, portable noise  -> all parity checks
#   fast   : -O3 -march=x86-64-v3 (no -ffast-math)                 -> reported CPU baseline (optional)
#   stock  : -O3 -ffast-math -march=x86-64-v3 + BUILD_X86 (the reference's AVX / AVX2-FMA3 noise, picked by cpuid at run time):
#            the flags unix/configure.ac:750-763,799 selects, with -march=native replaced by x86-64-v3 because the binary is built
#            in one container and timed on another host.  -ffast-math is applied to source/core (every TU the trace time is
#            spent in); with gcc 13 the parser does not survive -ffinite-math-only ("Cannot parse input" on any scene), so the
#            remaining TUs are built with -O3 -march=x86-64-v3 only        -> second reported CPU baseline
#
# usage: oracle/build_ref.sh [parity|fast|stock] [jobs]
set -euo pipefail
VARIANT=${1:-parity}
CORE_OPT=""
JOBS=${2:-$(nproc)}
R=${POV_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
B=$HERE/_ref/$VARIANT
if [ ! -d "$R/source" ]; then echo "reference sources not present at $R; using prebuilt $B"; exit 0; fi
mkdir -p "$B/obj"
case $VARIANT in
  parity) OPT="-O2 -fno-fast-math -ffp-contract=off";;
  fast)   OPT="-O3 -march=x86-64-v3 -fno-fast-math";;
  stock)  OPT="-O3 -march=x86-64-v3"; CORE_OPT="-ffast-math";;
  *) echo "unknown variant"; exit 2;;
esac
cat > "$B/config.h" <<CFG
#define VERSION_BASE "3.8"
#define PACKAGE "povray"
#define PACKAGE_NAME "POV-Ray"
#define PACKAGE_VERSION "3.8.0"
#define VERSION "3.8.0"
#define BUILT_BY "povray_b200 oracle build ($VARIANT)"
#define BUILD_ARCH "x86_64-pc-linux-gnu"
#define BUILT_FOR "x86_64-pc-linux-gnu"
#define COMPILER_VENDOR "gnu"
#define COMPILER_VERSION "g++"
#define POV_COMPILER_INFO "g++ @ x86_64-pc-linux-gnu"
#define CXXFLAGS "$OPT"
#define USE_OFFICIAL_BOOST
$( [ "$VARIANT" = stock ] && echo "#define BUILD_X86 1" )
#define LIBJPEG_MISSING
#define LIBTIFF_MISSING
#define OPENEXR_MISSING
#define X_DISPLAY_MISSING
#define IO_RESTRICTIONS_DISABLED 1
#define HAVE_NAN
#define HAVE_STD_ISNAN
#define HAVE_INF
#define HAVE_STD_ISINF
#define BUILTIN_IO_RESTRICTIONS "disabled"
#define BUILTIN_XWIN_DISPLAY "disabled"
#define BUILTIN_IMG_FORMATS "gif tga iff ppm pgm hdr png"
#define MISSING_IMG_FORMATS "jpeg tiff openexr"
#define POVLIBDIR "/nonexistent/share"
#define POVCONFDIR "/nonexistent/etc"
#define POVCONFDIR_BACKWARD "/nonexistent/etc"
#define HAVE_SYS_TIME_H 1
#define HAVE_SYS_RESOURCE_H 1
#define HAVE_SYS_WAIT_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_UNISTD_H 1
#define HAVE_TIME_H 1
#define HAVE_STDINT_H 1
#define HAVE_LIMITS_H 1
#define HAVE_GETRUSAGE 1
#define HAVE_GETTIMEOFDAY 1
#define HAVE_CLOCK_GETTIME 1
#define HAVE_NANOSLEEP 1
#define HAVE_USLEEP 1
#define HAVE_GETCWD 1
#define HAVE_READLINK 1
#define HAVE_SIGTIMEDWAIT 1
#define HAVE_USECONDS_T 1
#define HAVE_CLOCKID_T 1
#define HAVE_DECL_CLOCK_MONOTONIC 1
#define HAVE_DECL_CLOCK_REALTIME 1
#define HAVE_DECL_CLOCK_PROCESS_CPUTIME_ID 1
#define HAVE_DECL_CLOCK_THREAD_CPUTIME_ID 1
#define HAVE_DECL_RUSAGE_SELF 1
#define HAVE_DECL_RUSAGE_THREAD 1
CFG
cp "$R/libraries/png/scripts/pnglibconf.h.prebuilt" "$B/pnglibconf.h"
cat > "$B/jpeg_stub.cpp" <<STUB
#include "base/image/jpeg_pov.h"
#include "base/pov_err.h"
namespace pov_base { namespace Jpeg { Image* Read(IStream*, const ImageReadOptions&)
  { throw POV_EXCEPTION_STRING("jpeg disabled"); } }}
STUB
INC="-DHAVE_CONFIG_H -I$B -I$R/source -I$R/unix/povconfig -I$R/unix -I$R/vfe -I$R/vfe/unix -I$R/platform/unix -I$R/platform/x86 -I$R/libraries/boost -I$R/libraries/png -I$R/libraries/zlib"
CXXF="-std=c++11 $OPT -pthread -w -fPIC"
echo "$INC" > "$B/inc.flags"; echo "$CXXF" > "$B/cxx.flags"
# generate a makefile: one rule per TU, object names flattened
MK=$B/Makefile.gen
{
  echo "all: objs"
  OBJS=""
  while read -r f; do
    o=$B/obj/$(echo "${f#$R/}" | tr '/' '_' | sed 's/\.cpp$/.o/')
    extra=""
    case $f in */source/core/*) extra="${CORE_OPT:-}";; esac
    case $f in */avx/*) extra="-mavx";; */avxfma4/*) extra="-mavx -mfma4";; */avx2fma3/*) extra="-mavx2 -mfma";; esac
    echo "$o: $f"; printf '\tg++ %s %s %s -c $< -o $@\n' "$CXXF" "$extra" "$INC"
    OBJS="$OBJS $o"
  done < <(find "$R/source" "$R/vfe" "$R/platform/unix" "$R/platform/x86" -name '*.cpp' -not -path '*/vfe/win/*' | sort; echo "$R/unix/disp_text.cpp")
  o=$B/obj/jpeg_stub.o; echo "$o: $B/jpeg_stub.cpp"; printf '\tg++ %s %s -c $< -o $@\n' "$CXXF" "$INC"; OBJS="$OBJS $o"
  while read -r f; do
    o=$B/obj/$(echo "${f#$R/}" | tr '/' '_' | sed 's/\.c$/.o/')
    echo "$o: $f"; printf '\tgcc -O2 -w -fPIC %s -c $< -o $@\n' "$INC"
    OBJS="$OBJS $o"
  done < <(ls "$R"/libraries/zlib/*.c "$R"/libraries/png/*.c | grep -v -e example.c -e minigzip.c -e pngtest.c)
  echo "objs:$OBJS"
} > "$MK"
make -f "$MK" -j"$JOBS" objs
# povray binary, and an archive of everything but main() for the harness/adapter to link against
g++ -pthread -o "$B/povray" "$B"/obj/*.o -lrt
rm -f "$B/libpovref.a"
ar rcs "$B/libpovref.a" $(ls "$B"/obj/*.o | grep -v vfe_unix_unixconsole.o)
echo "built $B/povray"
