"""ORACLE — test infrastructure only.  CPU restatement of the reference's hot path (see vit_oracle.py,
vq_oracle.c).  The product package (vit-is-all-you-need_b200/) never imports anything from here."""
