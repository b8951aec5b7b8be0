# Executed inside the ``dim_b200`` namespace (see dim_b200/__init__.py).
__version__ = "0.1.0"
