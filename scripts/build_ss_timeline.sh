# development build of the library with the per-role wait counters of spconv_ss_kernel compiled in (-DSGB_SS_TIMELINE);
# never loaded by the product (scripts/ss_timeline.py points softgroup_b200.ops._lib.LIB_PATH at it)
set -e
mkdir -p scripts/experiments/build/tl
for f in softgroup_b200/csrc/*.cu; do
  o=scripts/experiments/build/tl/$(basename ${f%.cu}).o
  if [ ! -f $o ] || [ $f -nt $o ] || [ "$(basename $f)" = spconv_ss.cu ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DSGB_SS_TIMELINE -c $f -o $o &
  fi
done
wait
nvcc -shared -o scripts/experiments/build/libsgb200_tl.so scripts/experiments/build/tl/*.o -gencode arch=compute_100a,code=sm_100a -cudart static
ls -la scripts/experiments/build/libsgb200_tl.so
