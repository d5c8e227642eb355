# tools/build_variant.sh NAME [-DFLAG ...] : another build of the library into build/variants/NAME.so (tools/sweep_geom.py, RAPPAS_B200_LIB)
set -e
name=$1; shift
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2,-Wall,-fvisibility=hidden --expt-relaxed-constexpr \
  -fmad=false -prec-div=true -prec-sqrt=true "$@" -shared -o build/variants/$name.so \
  rappas_b200/csrc/rp_db.cu rappas_b200/csrc/rp_place.cu rappas_b200/csrc/rp_dbbuild.cu rappas_b200/csrc/rp_synthdb.cu rappas_b200/csrc/rp_xchg.cu rappas_b200/csrc/rp_ingest.cpp -lcudart -ldl
echo build/variants/$name.so
