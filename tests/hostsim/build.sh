#!/bin/bash
# builds the host simulator of the device code (test infrastructure)
set -e
cd "$(dirname "$0")"
g++ -O2 -std=c++17 -shared -fPIC -I../../parallel-in-time-ode-filters_b200/csrc -x c++ hostsim.cpp -o libhostsim.so
