mkdir -p gpurun_out
export ONLY=5 CAPS=0
timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python tools/diag_fp32.py > gpurun_out/san_memcheck_diag.log 2>&1
timeout 900 compute-sanitizer --tool initcheck --print-limit 30 python tools/diag_fp32.py > gpurun_out/san_initcheck_diag.log 2>&1
tail -40 gpurun_out/san_memcheck_diag.log; tail -60 gpurun_out/san_initcheck_diag.log
