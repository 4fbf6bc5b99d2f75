"""Command-line front end for the NPZ wire format of the reference's ``part-dens`` program."""
