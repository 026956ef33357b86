set -x
python tools/pcie_bw.py 0
python tools/pcie_bw.py 1
python tools/pcie_bw.py 0 & python tools/pcie_bw.py 1 & wait
