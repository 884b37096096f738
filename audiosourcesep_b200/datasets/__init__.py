"""Mirror of the reference's ``datasets`` package for the separation path (data_loader.get_song_extract)."""
