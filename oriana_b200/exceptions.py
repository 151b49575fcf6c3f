"""The reference's two exception types (oriana/exceptions.py:6-11)."""


class DatatypeException(Exception):
    """Raised when a count matrix is built from an unsupported container (cmatrix.py:25-29)."""


class IncompatibleShapeException(Exception):
    """Raised when a dimension relation string is malformed (dims.py:107-109)."""
