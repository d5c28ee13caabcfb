#!/bin/bash
# round 2, call 3o (1 GPU): per-CTA balance of the event phases of the fused kernel
timeout 300 python scratch/cta_balance.py 2>&1 | tail -14
