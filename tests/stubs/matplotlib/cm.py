def get_cmap(name):
    raise RuntimeError('matplotlib stub: colour maps are not available in the test harness')
