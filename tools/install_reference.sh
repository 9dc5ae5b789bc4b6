#!/bin/sh
# Offline install of the UNMODIFIED reference into the git-ignored baseline/_ref (bench.py --impl reference and the
# cpu_baseline leg time it; the directory travels to the GPU box with the gpurun snapshot).  Build container only:
# /root/reference does not exist on the GPU box.
set -e
cd "$(dirname "$0")/.."
rm -rf baseline/_ref
mkdir -p baseline
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref "${1:-/root/reference}"
