# mr_gen/model/utils/values.py:2 of the reference: the collate functions pad with this value
PADDING_VALUE = -100
