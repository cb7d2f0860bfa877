"""profiles/r2_ptxas_summary.csv: registers and spill bytes of every kernel from the `-Xptxas -v` logs the Makefile keeps
(csrc/*.ptxas.log).  python tools/ptxas_summary.py > profiles/r2_ptxas_summary.csv"""
import glob
import os
import re
import subprocess

root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nbodysimproject_b200", "csrc")
print("translation_unit,kernel,registers,spill_store_bytes,spill_load_bytes")
for path in sorted(glob.glob(os.path.join(root, "*.ptxas.log"))):
    tu = os.path.basename(path)[:-len(".ptxas.log")]
    name = None
    spill = (0, 0)
    for line in open(path):
        m = re.search(r"Compiling entry function '([^']+)'", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and name:
            spill = (int(m.group(1)), int(m.group(2)))
            continue
        m = re.search(r"Used (\d+) registers", line)
        if m and name:
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
            print(f'{tu},"{dem}",{m.group(1)},{spill[0]},{spill[1]}')
            name = None
