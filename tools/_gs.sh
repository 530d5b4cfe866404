mkdir -p gpurun_out/s12
tools/probes/bin/powf_probe | tee gpurun_out/s12/powf_probe.txt
