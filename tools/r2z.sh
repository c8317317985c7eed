MTX_NVCC_DEFINES=MTX_PK_EVENTS python -c "
from maxtext_indextts2_b200 import _lib
_lib.build(force=True)
" > gpurun_out/r2z_build.log 2>&1
timeout 600 python tools/mega_trace.py > gpurun_out/r2z_trace.txt 2>&1
tail -5 gpurun_out/r2z_trace.txt
