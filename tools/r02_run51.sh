#!/bin/bash
SPLITS=2 timeout 300 python tools/attn_fwd_trace.py 2>&1 | grep -v Warn | tail -12
