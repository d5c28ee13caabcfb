#!/bin/bash
# round 2, call 3c (1 GPU): kernel classes of small back-end windows
timeout 300 python scratch/be_small_classes.py 2>&1 | tail -4
