import numpy as np

SEMPY_SCALAR = np.float64
