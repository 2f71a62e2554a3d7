#!/bin/sh
# Offline install of the UNMODIFIED reference (facebookresearch/audiocraft @ 896ec7c) into baseline/_ref.
# Run in the build container (the only place /root/reference exists).  --no-deps: the reference's third-party
# dependencies (xformers, flashy, julius, av, ...) are absent from the wheelhouse; oracle/ref_import.py stubs the ones
# the import graph touches (the hot path itself never calls them).  baseline/_ref is git-ignored and NOT
# gpurun-ignored, so it travels to the GPU box.
set -e
cd "$(dirname "$0")/.."
rm -rf baseline/_ref
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target baseline/_ref /root/reference
