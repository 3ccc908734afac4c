"""Rendering/spline constants with the reference's attribute names (pybpl/parameters.py:19-63).
Only the attributes the render+score path reads are mirrored."""


class Parameters:
    def __init__(self):
        self.imsize = (105, 105)        # parameters.py:21 (any (H, W): other sizes run the run-time-geometry kernels)
        self.ink_pp = 2.                # :25
        self.ink_max_dist = 2.          # :27
        self.ink_ncon = 2               # :31
        self.ink_a = 0.5                # :33
        self.ink_b = 6.                 # :35
        self.broaden_mode = 'Lake'      # :37
        self.fsize = 11                 # :41
        self.spline_max_neval = 200     # :45
        self.spline_min_neval = 10      # :47
        self.spline_grain = 1.5         # :49
        self.max_blur_sigma = 16.       # :56
        self.min_blur_sigma = 0.5       # :58
        self.max_epsilon = 0.5          # :60
        self.min_epsilon = 1e-4         # :62
