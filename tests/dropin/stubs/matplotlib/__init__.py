"""TEST INFRASTRUCTURE: stand-in for `matplotlib` (absent from this image): the reference examples import pyplot at
module level and only use it when display / plotting is requested."""
