__version__ = "0.1.0+pylinac3.46.0"
