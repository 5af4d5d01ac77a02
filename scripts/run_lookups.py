#!/usr/bin/env python
# -*- coding: utf-8 -*-
''' Create lookup tables for specific neurons on the GPU (same flags as the reference's
    scripts/run_lookups.py).  Example: python scripts/run_lookups.py -n RS -a 32 -f 500 --mpi '''

import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from pysonic_b200.run_lookups import main  # noqa: E402

if __name__ == '__main__':
    main()
