"""Namespace package for the B200 backend of x42/phaserotate.lv2 (see phaserotate.lv2_b200)."""
