#!/bin/bash
# round 2, call m (1 GPU): kernel times of one rank's share of C5 / C4 sharded over 8, whole-plane vs row-band pieces
timeout 300 python scratch/be_band_probe.py C5 8 2>&1 | tail -8
timeout 300 python scratch/be_band_probe.py C4 8 2>&1 | tail -8
