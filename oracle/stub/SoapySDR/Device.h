/*
 * TEST INFRASTRUCTURE — stub of the SoapySDR C API, just enough for the reference's soapy.c to
 * compile in place (oracle/ref_soapy_harness.c).  SoapySDR is absent from this image.
 * Call sites: soapy.c:80-163, 178-222, 297-318.
 */
#ifndef ORACLE_STUB_SOAPY_DEVICE_H
#define ORACLE_STUB_SOAPY_DEVICE_H
#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>
typedef struct SoapySDRDevice SoapySDRDevice;
typedef struct SoapySDRStream SoapySDRStream;
#define SOAPY_SDR_TX 0
#define SOAPY_SDR_RX 1
SoapySDRDevice *SoapySDRDevice_makeStrArgs(const char *args);
int SoapySDRDevice_unmake(SoapySDRDevice *d);
const char *SoapySDRDevice_lastError(void);
int SoapySDRDevice_setGainMode(SoapySDRDevice *d, int dir, size_t ch, bool automatic);
int SoapySDRDevice_setGain(SoapySDRDevice *d, int dir, size_t ch, double value);
int SoapySDRDevice_setFrequencyCorrection(SoapySDRDevice *d, int dir, size_t ch, double value);
int SoapySDRDevice_setFrequency(SoapySDRDevice *d, int dir, size_t ch, double f, const void *args);
int SoapySDRDevice_setSampleRate(SoapySDRDevice *d, int dir, size_t ch, double rate);
int SoapySDRDevice_setAntenna(SoapySDRDevice *d, int dir, size_t ch, const char *name);
SoapySDRStream *SoapySDRDevice_setupStream(SoapySDRDevice *d, int dir, const char *format, const size_t *chans, size_t nchans, const void *args);
int SoapySDRDevice_closeStream(SoapySDRDevice *d, SoapySDRStream *s);
int SoapySDRDevice_activateStream(SoapySDRDevice *d, SoapySDRStream *s, int flags, long long timeNs, size_t numElems);
int SoapySDRDevice_deactivateStream(SoapySDRDevice *d, SoapySDRStream *s, int flags, long long timeNs);
int SoapySDRDevice_readStream(SoapySDRDevice *d, SoapySDRStream *s, void *const *buffs, size_t numElems, int *flags, long long *timeNs, long timeoutUs);
#endif
