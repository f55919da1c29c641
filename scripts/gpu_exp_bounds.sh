set -x
mkdir -p gpurun_out /tmp/try
for V in "4 5" "5 5" "5 6" "6 6" "0 0"; do set -- $V; SB=$1; PB=$2
  mkdir -p /tmp/try/$SB$PB/voxel-rt_b200/csrc /tmp/try/$SB$PB/include
  cp voxel-rt_b200/csrc/*.cu* /tmp/try/$SB$PB/voxel-rt_b200/csrc/; cp include/vxrt.h /tmp/try/$SB$PB/include/
  if [ "$SB" != "0" ]; then sed -i "s/__launch_bounds__(256) shade_kernel/__launch_bounds__(256, $SB) shade_kernel/; s/__launch_bounds__(256) primary_kernel/__launch_bounds__(256, $PB) primary_kernel/" /tmp/try/$SB$PB/voxel-rt_b200/csrc/kernels.cuh; fi
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -Xcompiler -fPIC -shared -o /tmp/try/lib_$SB$PB.so /tmp/try/$SB$PB/voxel-rt_b200/csrc/vxrt.cu
  echo "== shade minblocks $SB primary minblocks $PB"
  VXRT_LIB=/tmp/try/lib_$SB$PB.so python scripts/exp_time.py 2>&1 | tail -3
done
