import sys, numpy as np
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import parity
from porousfreezethaw_b200 import cases
parity.check_rhs(cases.small_case(24, 20, 32, calc_mode=0), None)
print("rhs ok")
