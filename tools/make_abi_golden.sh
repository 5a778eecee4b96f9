#!/bin/sh
# Regenerates tests/golden/cudpp_abi.txt from the reference's own cudpp.h (needs /root/reference).
set -e
cd "$(dirname "$0")/.."
g++ -I/root/reference/cudpp-inpar/include tests/c/cudpp_enum_dump.cc -o /tmp/cudpp_enum_ref
/tmp/cudpp_enum_ref > tests/golden/cudpp_abi.txt
wc -l tests/golden/cudpp_abi.txt
