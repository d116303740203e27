// file_driver.cpp -- test driver for the file route of the C++ facade (AudioFormatReader / AudioFilePlayer).
//   file_driver header <file>                      print the parsed container fields (no GPU needed)
//   file_driver analyse <file> <tracks> <out.f32>  load the file into the transport and analyse it on the GPU:
//                                                  writes [tracks][frames][12] smoothed features, prints "frames N"
#include "../../feature-extractor_b200/host/FeatureExtractorB200.h"

#include <cstdio>
#include <cstdlib>

using namespace fxb200;

int main (int argc, char** argv)
{
    if (argc < 3) { fprintf (stderr, "usage\n"); return 2; }
    const String mode = argv[1];
    if (mode == "header")
    {
        auto r = AudioFormatReader::createReaderFor (String (argv[2]));
        if (! r) { printf ("unreadable\n"); return 0; }
        printf ("%s rate %.3f channels %d bits %d float %d frames %ld format %d bytes %zu\n", r->formatName.c_str(), r->sampleRate,
                r->numChannels, r->bitsPerSample, r->usesFloatingPointData ? 1 : 0, r->lengthInSamples, r->pcmFormat, r->dataBytes());
        return 0;
    }
    if (mode == "analyse" && argc >= 5)
    {
        const int T = atoi (argv[3]);
        auto probe = AudioFormatReader::createReaderFor (String (argv[2]));
        if (! probe) { fprintf (stderr, "unreadable\n"); return 3; }
        AudioDeviceManager deviceManager (T, probe->sampleRate, 512, 2048);
        AudioFilePlayer player;
        player.setupAudioCallback (deviceManager);
        player.loadFileIntoTransport (String (argv[2]));
        if (! player.hasFile()) return 3;
        player.play();
        std::vector<float> smoothed;
        const long frames = player.analyseLoadedFile (smoothed);
        if (frames < 0) { fprintf (stderr, "analyse failed: %s\n", fx_last_error (deviceManager.getEngine())); return 4; }
        FILE* out = fopen (argv[4], "wb");
        fwrite (smoothed.data(), sizeof (float), smoothed.size(), out);
        fclose (out);
        printf ("frames %ld\n", frames);
        return 0;
    }
    return 2;
}
