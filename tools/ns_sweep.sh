for K in 65536 131072 262144 524288; do for ns in 1 2; do
echo "K=$K NS=$ns $(MPPI_NS=$ns python tools/profile_step.py --K $K --T 100 --steps 12 --timing 2>&1 | grep -o "'rollout': [0-9.]*")"
done; done
for E in 128 256 512; do for ns in 1 2; do
echo "C5 envs=$E NS=$ns $(MPPI_NS=$ns python tools/profile_batched.py --envs $E --steps 12 --timing 2>&1 | grep -o "'rollout': [0-9.]*")"
done; done
