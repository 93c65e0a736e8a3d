#!/bin/sh
# Recipe for the reference-as-shipped CPU arm (bench.py --impl reference, oracle/ref_arm.py):
# installs the UNMODIFIED reference package into git-ignored baseline/_ref (it travels to the GPU box with the gpurun snapshot).
# /root/reference is read-only, so the wheel is built from a /tmp copy; --no-deps because matplotlib / seaborn /
# vanilla-option-pricers are not in the offline wheelhouse (oracle/ref_arm.py stubs them; the timed calls never touch them).
set -e
cd "$(dirname "$0")/.."
rm -rf /tmp/b200sv_refcopy baseline/_ref
cp -r /root/reference /tmp/b200sv_refcopy
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref /tmp/b200sv_refcopy
rm -rf /tmp/b200sv_refcopy
