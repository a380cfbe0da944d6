#!/usr/bin/env python
"""Drop-in for the reference's matcher.py (same argv and files); implementation in pfann_b200/cli.py."""
import sys

from pfann_b200.cli import matcher_main

if __name__ == '__main__':
    sys.exit(matcher_main(sys.argv))
