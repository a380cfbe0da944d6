#!/usr/bin/env python
"""Drop-in for the reference's extractemb.py (same argv and files); implementation in pfann_b200/cli.py."""
import sys

from pfann_b200.cli import extractemb_main

if __name__ == '__main__':
    sys.exit(extractemb_main(sys.argv))
