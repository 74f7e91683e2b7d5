"""ORACLE — CPU restatement of the reference MPL forward. Test infrastructure only (see mpl_oracle.py)."""
