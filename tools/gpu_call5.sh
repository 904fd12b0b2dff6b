set -x
timeout 200 python tools/quant_phase.py 256 1280 > gpurun_out/c5_quant_phase_256x1280.txt 2>&1
timeout 200 python tools/quant_phase.py 1024 640 > gpurun_out/c5_quant_phase_1024x640.txt 2>&1
cat gpurun_out/c5_quant_phase_256x1280.txt | tail -14
