mkdir -p gpurun_out/s15
run () { for mnk in 8192,8192,8192:nn 8192,8192,8192:tn 4096,4096,4096:nn 2048,2048,2048:nn 784,1024,8192:tn 1024,1024,8192:tn 8192,8192,1024:nt; do python tools/gemm_bench.py --mnk ${mnk%%:*} --layouts ${mnk##*:} --iters 5 2>&1 | grep TFLOP; done; }
echo "== new (2 CTAs/SM for [k][row] operands)"; run
cp cortenn_b200/libllo_cuda.so /tmp/new.so; cp tools/_old_libllo_cuda.so cortenn_b200/libllo_cuda.so
echo "== old"; run
cp /tmp/new.so cortenn_b200/libllo_cuda.so
echo "== new again"; run
python tools/gemm_bench.py --check --size 512 --iters 1 2>&1 | tail -2
