set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_host.py tests/test_gpu_parity.py -m gpu -x -q -k "gl_shim or upload_rows" 2>&1 | tail -5
# the same binary as a 4K session: 300 frames after the depth threads finished, a destruction every 50 frames
d=$(mktemp -d); printf '#version 430\nvoid main(){}\n' > $d/vshader.glsl
printf '#version 430\nconst int VOXELS_WIDTH=512;\nconst int VOXELS_HEIGHT=96;\nconst int RENDER_DIST=384;\nvoid main(){}\n' > $d/fshader.glsl
(cd $d && VXRT_GLSHIM_READY_UPLOADS=2 VXRT_GLSHIM_FRAMES=300 VXRT_GLSHIM_FPS=60 VXRT_GLSHIM_LOG=1 \
  VXRT_GLSHIM_EVENTS="0:resize:3840x2160;1:mouse:400,600;1:lmb:down;9:lmb:up;50:rmb:down;51:rmb:up;100:rmb:down;101:rmb:up" \
  VXRT_GLSHIM_DUMP=$GRAFT_REPO_ROOT/gpurun_out/glshim_%03d.ppm VXRT_GLSHIM_DUMP_FRAMES=299 \
  $GRAFT_REPO_ROOT/oracle/_ref/voxel_rt_on_vxrt) 2>&1 | tail -3 | tee gpurun_out/glshim_session.txt
