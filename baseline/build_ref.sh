#!/bin/bash
# Builds the UNMODIFIED reference CUDA extension (`cuam`) for sm_100 as a reported baseline and as the
# generator of tests/golden/ref_*.json.  Recipe from SURVEY.md App. C (the reference's CMake file has no
# CUDA arch and asks for C++14, which torch 2.11 rejects) + a link-time guard for zero-sized launches
# (baseline/zero_launch_guard.cpp).  Output: baseline/_ref/AnalyticMesh/backend/build/cuam.so
# (git-ignored; travels to the GPU box with the gpurun snapshot).  Takes ~6 minutes on 8 cores.
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OBJ=${OBJ:-/tmp/refbuild}
DST=$HERE/_ref/AnalyticMesh
mkdir -p "$OBJ" "$DST"
if [ ! -d "$DST/backend" ]; then
  cp -r "$REF/backend" "$REF/__init__.py" "$REF/examples" "$REF/LICENSE" "$DST/"
  mkdir -p "$DST/backend/build/libpolytools"
  touch "$DST/backend/build/__init__.py" "$DST/backend/build/libpolytools/__init__.py"
fi
TINC=$(python -c "import torch,os;print(os.path.join(os.path.dirname(torch.__file__),'include'))")
TLIB=$(dirname "$TINC")/lib
PYINC=$(python -c "import sysconfig;print(sysconfig.get_paths()['include'])")
F="-std=c++17 -O2 -rdc=true -gencode arch=compute_100,code=sm_100 --expt-relaxed-constexpr -Xcompiler -fPIC -DTORCH_EXTENSION_NAME=cuam -I$REF/backend/inc -I$TINC -I$TINC/torch/csrc/api/include -I$PYINC"
cd "$OBJ"
[ -f cuam_kernel.o ] || nvcc $F -c "$REF/backend/src/cuam_kernel.cu" -o cuam_kernel.o &
[ -f kernel.o ]      || nvcc $F -c "$REF/backend/src/kernel.cu" -o kernel.o &
[ -f states.o ]      || nvcc $F -c "$REF/backend/src/states.cu" -o states.o &
[ -f utilities.o ]   || nvcc $F -x cu -c "$REF/backend/src/utilities.cpp" -o utilities.o &
[ -f cuam.o ]        || /usr/bin/g++ -std=c++17 -O2 -fPIC -DTORCH_EXTENSION_NAME=cuam -I"$REF/backend/inc" -I"$TINC" -I"$TINC/torch/csrc/api/include" -I"$PYINC" -I/usr/local/cuda/include -c "$REF/backend/src/cuam.cpp" -o cuam.o &
wait
/usr/bin/g++ -O2 -fPIC -I/usr/local/cuda/include -c "$HERE/zero_launch_guard.cpp" -o zero_launch_guard.o
nvcc -shared -rdc=true -gencode arch=compute_100,code=sm_100 -Xcompiler -fPIC -Xlinker --wrap=cudaLaunchKernel \
     -o "$DST/backend/build/cuam.so" cuam_kernel.o kernel.o states.o utilities.o cuam.o zero_launch_guard.o \
     -L"$TLIB" -ltorch_python -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -lcublas -lcudadevrt
echo "built $DST/backend/build/cuam.so"
