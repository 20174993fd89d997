"""Import-only stand-in (see vkit/element.py)."""
stack_flattened_text_regions = FlattenedTextRegion = TextRegionFlattener = None

