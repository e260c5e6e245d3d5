#!/bin/bash
timeout 900 python tools/big_check.py 2>&1 | tail -8
