export QG_DEAD=trace
QG_KERNELS=mega python tools/quick_grid.py 2>&1 | tail -1
QG_KERNELS=grid_stream python tools/quick_grid.py 2>&1 | tail -1
export PT_STREAM_ALL=1
for th in 4 12 16 20 24 28; do for b in 1 6; do echo "thresh $th batch $b"; PT_STREAM=$((th + b*256)) QG_KERNELS=grid_stream python tools/quick_grid.py 2>&1 | tail -1; done; done
