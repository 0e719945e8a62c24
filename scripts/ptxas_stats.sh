#!/bin/bash
# registers / spills per kernel of one source file: scripts/ptxas_stats.sh stencil3d.cu [filter]
cd "$(dirname "$0")/../pde_solver_b200/csrc" || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
  $PDE_B200_NVCC_FLAGS -Xptxas -v -x cu -c "$1" -o /tmp/ptxas_stats.o 2>&1 | python3 -c "
import sys,re,subprocess
name=None; spill=''
for l in sys.stdin:
    m=re.search(r\"Compiling entry function '(\S+)'\",l)
    if m: name=m.group(1)
    if 'spill stores' in l: spill=l.strip()
    m=re.search(r'Used (\d+) registers',l)
    if m and name:
        d=subprocess.run(['c++filt',name],stdout=subprocess.PIPE,text=True).stdout.split('(')[0]
        print(f'{d:50s} regs {m.group(1):>4s}  {spill}')
" | grep -E "${2:-.}"
