set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:tap_gemm -s 60 -c 4 -o gpurun_out/r02_gemm_epi -f python tools/codec_breakdown.py > gpurun_out/ncu_gemm_epi.log 2>&1; tail -2 gpurun_out/ncu_gemm_epi.log
ls -la gpurun_out/r02_gemm_epi.ncu-rep
