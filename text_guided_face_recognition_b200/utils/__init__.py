"""Mirror of the scoring half of the reference's `utils/modules.py`."""
