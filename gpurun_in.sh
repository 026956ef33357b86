set -x
python tools/pcie_bw.py
for c in 16 32 64; do CUHE_B200_HOST_CHUNK=$c CUHE_B200_HOST_RAMP=1 python tools/e2e_sweep.py --one 256; done
CUHE_B200_HOST_CHUNK=32 CUHE_B200_HOST_RAMP=0 python tools/e2e_sweep.py --one 256
