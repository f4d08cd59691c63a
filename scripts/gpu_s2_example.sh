#!/bin/bash
# the two example programs against the C ABI on a real GPU
g++ -std=c++17 -O1 -Iinclude examples/simple.cpp -Lmetrovector_b200 -l:libmvf_b200.so -Wl,-rpath,$PWD/metrovector_b200 -o /tmp/simple && /tmp/simple | tail -9
