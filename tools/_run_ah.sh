for ah in 1 2 3; do for cfg in "512 4000" "256 16000" "128 50000"; do echo -n "ahead=$ah "; MB200_LL_AHEAD=$ah python tools/fused_time.py $cfg 4 | sort -t' ' -k3 -n | head -1; done; done
