#!/bin/bash
# Quick look at the SASS of the two headline kernels (BASELINE configs[0] AC, configs[1] WM) without building the
# whole library: compiles scan_packed.cu with -DACWM_DEV_ONLY to a cubin under /tmp and prints resource usage.
set -e
cd "$(dirname "$0")/../cuda-aho-corasick-wu-manber_b200"
mkdir -p /tmp/sass
nvcc -std=c++17 -O3 -lineinfo --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -diag-suppress 186,177 \
	-DACWM_DEV_ONLY -cubin -o /tmp/sass/dev.cubin csrc/scan_packed.cu
cuobjdump -res-usage /tmp/sass/dev.cubin | grep -E "Function|REG" | paste - - | sed -E 's/.*scan_kernelINS_7(Front[A-Z]+).*REG:([0-9]+).*/\1 regs \2/'
for k in AC WM; do
	cuobjdump -sass /tmp/sass/dev.cubin | awk -v k="Front$k" '/Function :/ {on = index($0, k) > 0} on' | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\/$//' > /tmp/sass/dev_$k.txt
	echo "$k: $(wc -l < /tmp/sass/dev_$k.txt) SASS lines -> /tmp/sass/dev_$k.txt"
done
