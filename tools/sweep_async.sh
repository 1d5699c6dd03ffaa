export QG_DEAD=elide QG_KERNELS=grid_async
for k in 8 16 24 28 31 32; do for b in 1 4 16; do echo -n "K $k batch $b: "; PT_ASYNC=$((k + b*256)) python tools/quick_grid.py 2>&1 | tail -1 | cut -c30-75; done; done
