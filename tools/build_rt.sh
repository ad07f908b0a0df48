# instrumented / alternative builds of the library next to the product build:
#   tools/build_rt.sh                 -> tools/libmicloc_b200_rt.so   (role timing + wait debug)
#   tools/build_rt.sh san             -> tools/libmicloc_b200_san.so  (-DMICLOC_UNALIGNED_BARRIERS: the synccheck-clean barrier form)
# use with MICLOC_B200_LIB=<path> (haghighatshoarmuir2024_b200/_native.py)
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=$HERE/../haghighatshoarmuir2024_b200/csrc
KIND=${1:-rt}
if [ "$KIND" = san ]; then DEFS="-DMICLOC_UNALIGNED_BARRIERS"; else DEFS="-DMICLOC_ROLE_TIMING -DMICLOC_WAIT_DEBUG"; fi
TMP=$(mktemp -d)
for f in api fused fused_tc rzcc xylo peak synth stream multiband; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr \
       $DEFS $EXTRA -I$SRC -c -o $TMP/$f.o $SRC/micloc_$f.cu 2> $TMP/$f.log &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o $HERE/libmicloc_b200_$KIND.so $TMP/*.o -lcudart_static -lpthread -ldl -lrt
rm -rf $TMP
