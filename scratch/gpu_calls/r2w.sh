#!/bin/bash
# round 2, call w (1 GPU): packet preparation / evaluation alone and under concurrent copies
timeout 300 python scratch/prep_under_copy.py 2>&1 | tail -16
