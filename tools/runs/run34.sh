python tools/tune_trace.py mesh1m '{"PTAP_VOTE_TRI": ["6", "8", "12"], "PTAP_VOTE_INST": ["4", "6", "10"], "PTAP_VOTE_REFILL": ["8"], "PTAP_BATCH": ["32"]}' 2>&1 | tail -12
python tools/tune_trace.py mesh1m '{"PTAP_VOTE_TRI": ["8"], "PTAP_VOTE_INST": ["6"], "PTAP_VOTE_REFILL": ["4", "16"], "PTAP_BATCH": ["32", "64"]}' 2>&1 | tail -8
