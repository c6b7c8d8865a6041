# Same-box A/B of two builds of libfs2d_cuda.so (tools/pcg_probe.py): current vs libfs2d_cuda_prev.so
cd flipsolver2d_b200; cp libfs2d_cuda.so libfs2d_cuda_new.so; cd ..
for rep in 1 2; do for v in new prev; do
  cp flipsolver2d_b200/libfs2d_cuda_$v.so flipsolver2d_b200/libfs2d_cuda.so
  echo "== $v"
  FS2D_PROBE_MODES=1 FS2D_PROBE_SOLVES=5 python tools/pcg_probe.py 4096 2>&1 | grep "^res"
done; done
cp flipsolver2d_b200/libfs2d_cuda_new.so flipsolver2d_b200/libfs2d_cuda.so
