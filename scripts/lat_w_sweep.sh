#!/bin/bash
for w in 1 2 3 4 5 6 8; do echo "W=$w"; FO_TEAM_WARPS=$w python scripts/latency_probe.py 2>/dev/null | head -2; done
