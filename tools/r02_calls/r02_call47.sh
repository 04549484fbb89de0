#!/bin/bash
timeout 100 python tools/video_windows_probe.py 272 32 2>&1 | tail -3
