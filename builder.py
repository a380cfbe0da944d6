#!/usr/bin/env python
"""Drop-in for the reference's builder.py (same argv and files); implementation in pfann_b200/cli.py."""
import sys

from pfann_b200.cli import builder_main

if __name__ == '__main__':
    sys.exit(builder_main(sys.argv))
