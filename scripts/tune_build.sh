# builds SpMM tuning variants of the library into gpurun_in/ (travels to the GPU box; *.so is git-ignored)
set -e
mkdir -p gpurun_in
cd fitgnn_b200/csrc
for v in "3 2 8" "4 2 8" "3 4 8" "2 4 8" "3 2 16" "4 2 16" "3 1 8" "4 1 8"; do
  set -- $v
  tag="m$1_u$2_r$3"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DFG_SPMM_MINB=$1 -DFG_SPMM_UNROLL=$2 -DFG_SPMM_RPW=$3 -c spmm.cu -o build/spmm_$tag.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../gpurun_in/libfitgnn_$tag.so build/capi.o build/primitives.o build/spmm_$tag.o build/gemm_simt.o build/gemm_tcgen05.o build/pool_project.o build/builders.o
done
ls -la ../../gpurun_in
