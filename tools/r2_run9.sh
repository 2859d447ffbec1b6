#!/bin/bash
# GPU session 9: full parity suite, then ncu captures of the kernels added in this session (encrypt_tab4, adder_chain_wide).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest9.txt
cat gpurun_out/r2_pytest9.txt
summarise() { ncu -i $1 --page details > $2_details.txt 2>/dev/null; ncu -i $1 --page raw --csv > $2_raw.csv 2>/dev/null; }
python tools/r2_kernel_zoo.py > gpurun_out/r2_zoo2_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"encrypt_tab4" -c 4 -o /tmp/r02_enc4 python tools/r2_kernel_zoo.py > gpurun_out/r2_zoo2_ncu.log 2>&1
tail -2 gpurun_out/r2_zoo2_ncu.log
summarise /tmp/r02_enc4.ncu-rep gpurun_out/r02_enc4
python tools/r2_kernel_zoo.py B > gpurun_out/r2_zoob2_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"adder_chain_wide" -c 1 -o /tmp/r02_wide python tools/r2_kernel_zoo.py B > gpurun_out/r2_zoob2_ncu.log 2>&1
tail -3 gpurun_out/r2_zoob2_plain.log; tail -2 gpurun_out/r2_zoob2_ncu.log
summarise /tmp/r02_wide.ncu-rep gpurun_out/r02_wide
ls -la gpurun_out/r02_enc4* gpurun_out/r02_wide*
