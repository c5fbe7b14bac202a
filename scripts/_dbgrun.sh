timeout 90 python scripts/pair_check.py 2>&1 | tail -11
RL_TC_TRACE=1 timeout 60 python scripts/pair_profile.py 2>&1 | tail -6
