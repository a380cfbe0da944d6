#!/usr/bin/env python
"""Drop-in for the reference's matchemb.py (same argv and files); implementation in pfann_b200/cli.py."""
import sys

from pfann_b200.cli import matchemb_main

if __name__ == '__main__':
    sys.exit(matchemb_main(sys.argv))
