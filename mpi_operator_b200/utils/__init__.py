"""Observability helpers: trace export, roofline reports."""
