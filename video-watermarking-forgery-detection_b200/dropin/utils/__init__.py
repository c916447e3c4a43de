"""Drop-in overlay: only utils/JPEG.py is replaced; the reference's other utils stay its own."""
