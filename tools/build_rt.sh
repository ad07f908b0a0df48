# instrumented build of the library (role timing), next to the product build: tools/libmicloc_b200_rt.so
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=$HERE/../haghighatshoarmuir2024_b200/csrc
TMP=$(mktemp -d)
for f in api fused fused_tc rzcc xylo peak synth stream multiband; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr \
       -DMICLOC_ROLE_TIMING -DMICLOC_WAIT_DEBUG $EXTRA -I$SRC -c -o $TMP/$f.o $SRC/micloc_$f.cu &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o $HERE/libmicloc_b200_rt.so $TMP/*.o -lcudart_static -lpthread -ldl -lrt
rm -rf $TMP
