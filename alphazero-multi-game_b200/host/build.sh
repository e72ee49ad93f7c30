#!/usr/bin/env bash
# Build the C++ host mirror + pybind module `_alphazero_cpp` (links libaz_b200.so; rpath $ORIGIN).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
PKG="$HERE/.."
PY="${PYTHON:-python}"
PYINC="$($PY -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
PBINC="$($PY -c 'import pybind11; print(pybind11.get_include())')"
SUFFIX="$($PY -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
JSONINC="$($PY -c 'import os, sysconfig; print(os.path.join(sysconfig.get_paths()["purelib"], "include", "cudnn_frontend", "thirdparty"))')"
CXX="${AZ_HOST_CXX:-/usr/bin/g++}"   # NOT $CXX: the image's /opt/gcc wrapper links a second, static libstdc++ into the module
$CXX -std=c++17 -O2 -fPIC -shared -ffp-contract=off -fvisibility=hidden -I"$PYINC" -I"$PBINC" -I"$JSONINC" -I/usr/local/cuda/include \
    "$HERE/alphazero_host.cpp" "$HERE/python_module.cpp" -o "$PKG/_alphazero_cpp$SUFFIX" \
    -L"$PKG" -laz_b200 -ldl -pthread -Wl,-rpath,'$ORIGIN'
echo "built $PKG/_alphazero_cpp$SUFFIX"
# the reference's `self_play` command over the same host classes
$CXX -std=c++17 -O2 -ffp-contract=off -I"$JSONINC" -I/usr/local/cuda/include "$HERE/alphazero_host.cpp" "$HERE/selfplay_main.cpp" -o "$PKG/self_play" \
    -L"$PKG" -laz_b200 -ldl -pthread -Wl,-rpath,'$ORIGIN'
echo "built $PKG/self_play"
